// eloc.cu -- one-pass sample-space local energy.
//
// Fuses what the reference does in three native calls plus ~6 torch passes over [n, M]
// (vmc/energy/eloc.py:326-397: get_comb_hij_fused -> WavefunctionLUT.lookup -> scatter ->
// divide -> multiply -> sum): for every sample the CTA enumerates the connected determinants,
// probes the hash index of the sorted unique-sample table, and only for the hits evaluates
// <x|H|x'> and accumulates (psi(x')/psi(x)) * H.  Nothing of size [n, M] ever reaches HBM.
//
// Determinism: fixed row -> thread mapping, shuffle tree, ordered sum over warps and splits;
// no atomics.  Per element the arithmetic is the reference's (psi'/psi0 first, then * H);
// only the order of the final sum differs from torch's reduction (tolerance 1e-12 rel).
#include "lut.cuh"

namespace pynqs {

constexpr int kElocThreads = 256;

struct Cplx {
  double re, im;
};

// numpy / c10 complex division (torch/headeronly/util/complex.h operator/=)
__device__ __forceinline__ Cplx cdiv(Cplx x, Cplx y) {
  const double a = x.re, b = x.im, c = y.re, d = y.im;
  const double ac = fabs(c), ad = fabs(d);
  Cplx r;
  if (ac >= ad) {
    if (ac == 0.0 && ad == 0.0) {
      r.re = a / ac;
      r.im = b / ad;
    } else {
      const double rat = d / c, scl = 1.0 / (c + d * rat);
      r.re = (a + b * rat) * scl;
      r.im = (b - a * rat) * scl;
    }
  } else {
    const double rat = c / d, scl = 1.0 / (d + c * rat);
    r.re = (a * rat + b) * scl;
    r.im = (b * rat - a) * scl;
  }
  return r;
}

template <bool CPLX>
__device__ __forceinline__ Cplx load_psi(const double *__restrict__ psi, long long id) {
  Cplx v;
  if (CPLX) {
    const double2 t = __ldg(reinterpret_cast<const double2 *>(psi) + id);
    v.re = t.x;
    v.im = t.y;
  } else {
    v.re = __ldg(psi + id);
    v.im = 0.0;
  }
  return v;
}

// ratio psi'/psi0 times a real H, accumulated
template <bool CPLX>
__device__ __forceinline__ void accumulate(Cplx &acc, Cplx pm, Cplx p0, double h) {
  if (CPLX) {
    const Cplx q = cdiv(pm, p0);
    acc.re += q.re * h;
    acc.im += q.im * h;
  } else {
    acc.re += (pm.re / p0.re) * h;
  }
}

template <int L, bool CPLX>
__global__ void __launch_bounds__(kElocThreads)
eloc_kernel(const u64 *__restrict__ bra, long long n, const double *__restrict__ h1e, const double *__restrict__ h2e,
            const u64 *__restrict__ key, const double *__restrict__ psi, long long N, const HashHeader *__restrict__ hdr,
            const double *__restrict__ hii, double *__restrict__ partial, double *__restrict__ psi0_out, int splits,
            ExcGeom g) {
  __shared__ OrbLists lists;
  __shared__ Cplx s_psi0;
  __shared__ Cplx s_warp[kElocThreads / 32];
  const long long s = blockIdx.x / splits;
  const int split = blockIdx.x - (int)(s * splits);
  if (s >= n) return;
  const Onv<L> x = load_onv<L>(bra + s * L);
  if (threadIdx.x < 32) build_lists<L>(x, g.sorb, g.noA, g.noB, lists, threadIdx.x);
  if (threadIdx.x == 32) {
    const long long id = hashed_search<L>(key, N, hdr, x);
    Cplx p0 = {0.0, 0.0};
    if (id >= 0) p0 = load_psi<CPLX>(psi, id);
    s_psi0 = p0;
    if (split == 0) {
      psi0_out[CPLX ? 2 * s : s] = p0.re;
      if (CPLX) psi0_out[2 * s + 1] = p0.im;
    }
  }
  __syncthreads();
  const Cplx p0 = s_psi0;

  Cplx acc = {0.0, 0.0};
  if (split == 0 && threadIdx.x == 0) accumulate<CPLX>(acc, p0, p0, hii[s]);  // row 0: (psi0/psi0) * H_xx

  const int chunk = (g.nsd + splits - 1) / splits;
  const int r_begin = split * chunk;
  const int r_end = min(g.nsd, r_begin + chunk);
  for (int r = r_begin + threadIdx.x; r < r_end; r += kElocThreads) {
    const Exc e = decode_exc(g, lists, r);
    const Onv<L> y = apply_exc<L>(x, e);
    const long long id = hashed_search<L>(key, N, hdr, y);
    if (id >= 0) {
      const double h = exc_element<L, double>(x, e, h1e, h2e, g.sorb);
      accumulate<CPLX>(acc, load_psi<CPLX>(psi, id), p0, h);
    }
  }

  // deterministic block reduction
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc.re += __shfl_xor_sync(0xffffffffu, acc.re, o);
    if (CPLX) acc.im += __shfl_xor_sync(0xffffffffu, acc.im, o);
  }
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    Cplx t = s_warp[0];
#pragma unroll
    for (int w = 1; w < kElocThreads / 32; ++w) {
      t.re += s_warp[w].re;
      t.im += s_warp[w].im;
    }
    const long long o = s * splits + split;
    if (CPLX) {
      partial[2 * o] = t.re;
      partial[2 * o + 1] = t.im;
    } else {
      partial[o] = t.re;
    }
  }
}

template <bool CPLX>
__global__ void eloc_finish_kernel(const double *__restrict__ partial, double *__restrict__ eloc, long long n, int splits) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  double re = 0.0, im = 0.0;
  for (int k = 0; k < splits; ++k) {
    if (CPLX) {
      re += partial[2 * (s * splits + k)];
      im += partial[2 * (s * splits + k) + 1];
    } else {
      re += partial[s * splits + k];
    }
  }
  if (CPLX) {
    eloc[2 * s] = re;
    eloc[2 * s + 1] = im;
  } else {
    eloc[s] = re;
  }
}

int eloc_splits(long long n, int nsd) {
  if (n <= 0) return 1;
  long long want = (148LL * 8 + n - 1) / n;
  long long cap = (nsd + 2047) / 2048;
  if (cap < 1) cap = 1;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

long long eloc_scratch_bytes(long long n, int nsd, int cplx) {
  const int splits = eloc_splits(n, nsd);
  const long long w = cplx ? 2 : 1;
  // hii[n] | partial[n * splits * w]
  return 8 * (n + n * splits * w) + 256;
}

int launch_diag_f64(const u64 *bra, const double *h1e, const double *h2e, double *out, long long n, long long stride, int L,
                    int sorb, int nele, cudaStream_t st);

template <int L>
static int launch_eloc_L(const u64 *bra, long long n, const double *h1e, const double *h2e, const u64 *key, const double *psi,
                         int cplx, long long N, const HashHeader *hdr, double *hii, double *partial, double *eloc,
                         double *psi0, int splits, const ExcGeom &g, cudaStream_t st) {
  const long long blocks = n * splits;
  if (blocks > 0x7fffffffLL) {
    set_error("eloc: n * splits = %lld exceeds the grid limit; split the batch", blocks);
    return 1;
  }
  double *dst = splits == 1 ? eloc : partial;
  if (cplx)
    eloc_kernel<L, true><<<(unsigned)blocks, kElocThreads, 0, st>>>(bra, n, h1e, h2e, key, psi, N, hdr, hii, dst, psi0, splits, g);
  else
    eloc_kernel<L, false><<<(unsigned)blocks, kElocThreads, 0, st>>>(bra, n, h1e, h2e, key, psi, N, hdr, hii, dst, psi0, splits, g);
  count_launch();
  if (int rc = check_launch("eloc_kernel")) return rc;
  if (splits > 1) {
    const unsigned fb = (unsigned)((n + 255) / 256);
    if (cplx) eloc_finish_kernel<true><<<fb, 256, 0, st>>>(partial, eloc, n, splits);
    else eloc_finish_kernel<false><<<fb, 256, 0, st>>>(partial, eloc, n, splits);
    count_launch();
    if (int rc = check_launch("eloc_finish_kernel")) return rc;
  }
  return 0;
}

int launch_eloc(const u64 *bra, long long n, const double *h1e, const double *h2e, const u64 *key, const double *psi, int cplx,
                long long N, const void *hash_ws, void *scratch, long long scratch_bytes, double *eloc, double *psi0,
                const ExcGeom &g, cudaStream_t st) {
  if (n == 0) return 0;
  const long long need = eloc_scratch_bytes(n, g.nsd, cplx);
  if (scratch_bytes < need) {
    set_error("eloc scratch too small: %lld < %lld bytes", scratch_bytes, need);
    return 4;
  }
  const int splits = eloc_splits(n, g.nsd);
  double *hii = reinterpret_cast<double *>(scratch);
  double *partial = hii + n;
  if (int rc = launch_diag_f64(bra, h1e, h2e, hii, n, 1, g.L, g.sorb, g.nele, st)) return rc;
  const HashHeader *hdr = reinterpret_cast<const HashHeader *>(hash_ws);
  switch (g.L) {
    case 1: return launch_eloc_L<1>(bra, n, h1e, h2e, key, psi, cplx, N, hdr, hii, partial, eloc, psi0, splits, g, st);
    case 2: return launch_eloc_L<2>(bra, n, h1e, h2e, key, psi, cplx, N, hdr, hii, partial, eloc, psi0, splits, g, st);
    case 3: return launch_eloc_L<3>(bra, n, h1e, h2e, key, psi, cplx, N, hdr, hii, partial, eloc, psi0, splits, g, st);
  }
  set_error("unsupported ONV length L=%d", g.L);
  return 1;
}

}  // namespace pynqs
