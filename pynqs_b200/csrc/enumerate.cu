// enumerate.cu -- connected-determinant enumeration, optionally fused with <x|H|x'>.
//
// Replaces the reference's K1 (get_merged_ovlst_kernel, cuda/kernel.cu:147-166), K2
// (get_comb_SD_kernel, :195-222) and K5 (get_comb_SD_fused_kernel, :224-277) plus the
// `repeat` pre-fill of cuda_tensor.cpp:239.  Differences in design, not in results:
//   * no merged[n, sorb] tensor in HBM: one warp derives the occupied/virtual lists from
//     popcounts of the bra words and keeps them in shared memory;
//   * every output byte is written exactly once, as 16-byte vectors: a thread owns two
//     consecutive rows (flat row index even), so comb stores are ulonglong2 and Hmat stores
//     are double2;
//   * 64-bit flat indexing (the reference overflows int beyond 2^31 elements);
//   * the diagonal <x|H|x> runs in its own thread-per-sample kernel, so no lane of the
//     enumeration warps serialises nele^2/2 gathers.
#include "common.cuh"

namespace pynqs {

constexpr int kEnumThreads = 256;
constexpr int kEnumTileRows = 4096;  // rows of one sample handled by one CTA

template <int L>
__device__ __forceinline__ void store_pair_rows(u64 *dst, const Onv<L> &r0, const Onv<L> &r1) {
  // dst is 16-byte aligned (flat row index even); 2L words contiguous
  u64 buf[2 * L];
#pragma unroll
  for (int i = 0; i < L; ++i) {
    buf[i] = r0.w[i];
    buf[L + i] = r1.w[i];
  }
  ulonglong2 *d2 = reinterpret_cast<ulonglong2 *>(dst);
#pragma unroll
  for (int i = 0; i < L; ++i) d2[i] = make_ulonglong2(buf[2 * i], buf[2 * i + 1]);
}

template <int L>
__device__ __forceinline__ void store_row(u64 *dst, const Onv<L> &r) {
#pragma unroll
  for (int i = 0; i < L; ++i) dst[i] = r.w[i];
}

template <int L, typename T, bool WITH_H>
__global__ void __launch_bounds__(kEnumThreads)
enumerate_kernel(const u64 *__restrict__ bra, const T *__restrict__ h1e, const T *__restrict__ h2e, u64 *__restrict__ comb,
                 T *__restrict__ hmat, long long n, int tiles_per_sample, ExcGeom g) {
  __shared__ OrbLists lists;
  const long long s = blockIdx.x / tiles_per_sample;
  const int tile = blockIdx.x - (int)(s * tiles_per_sample);
  if (s >= n) return;
  const Onv<L> x = load_onv<L>(bra + s * L);
  if (threadIdx.x < 32) build_lists<L>(x, g.sorb, g.noA, g.noB, lists, threadIdx.x);
  __syncthreads();

  const long long M = (long long)g.nsd + 1;
  const long long base = s * M;           // flat index of row 0 of this sample
  const int odd = (int)(base & 1);        // rows are paired so that base + m is even
  const long long m_lo = (long long)tile * kEnumTileRows - odd;
  for (int u = threadIdx.x; u < kEnumTileRows / 2; u += kEnumThreads) {
    const long long m0 = m_lo + 2 * u;  // base + m0 is even
    if (m0 >= M) break;
    Onv<L> row[2];
    T val[2];
    bool ok[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const long long m = m0 + t;
      ok[t] = (m >= 0) && (m < M);
      row[t] = x;
      val[t] = (T)0.0;
      if (ok[t] && m > 0) {
        const Exc e = decode_exc(g, lists, (int)(m - 1));
        row[t] = apply_exc<L>(x, e);
        if (WITH_H) val[t] = exc_element<L, T>(x, e, h1e, h2e, g.sorb);
      }
    }
    if (ok[0] && ok[1]) {
      store_pair_rows<L>(comb + (base + m0) * L, row[0], row[1]);
      if (WITH_H) {
        if (m0 == 0) {
          hmat[base + 1] = val[1];  // row 0 belongs to the diagonal kernel
        } else if (sizeof(T) == 8) {
          *reinterpret_cast<double2 *>(hmat + base + m0) = make_double2((double)val[0], (double)val[1]);
        } else {
          *reinterpret_cast<float2 *>(hmat + base + m0) = make_float2((float)val[0], (float)val[1]);
        }
      }
    } else {
#pragma unroll
      for (int t = 0; t < 2; ++t)
        if (ok[t]) {
          store_row<L>(comb + (base + m0 + t) * L, row[t]);
          if (WITH_H && (m0 + t) > 0) hmat[base + m0 + t] = val[t];
        }
    }
  }
}

// thread-per-sample diagonal: hmat[s*M] = <x|H|x>
template <int L, typename T>
__global__ void __launch_bounds__(128)
diag_kernel(const u64 *__restrict__ bra, const T *__restrict__ h1e, const T *__restrict__ h2e, T *__restrict__ out,
            long long n, long long stride, int sorb, int nele) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const Onv<L> x = load_onv<L>(bra + s * L);
  out[s * stride] = diag_element<L, T>(x, h1e, h2e, sorb, nele);
}

// flag_bit=True companion of get_comb_tensor: states[n, M, sorb] = +-1 of every comb row
// (cpu_tensor.cpp:186-190, excitation.cpp:171-181)
template <int L>
__global__ void __launch_bounds__(256)
states_kernel(const u64 *__restrict__ comb, double *__restrict__ states, long long rows, int sorb) {
  const long long total = rows * sorb;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / sorb;
    const int k = (int)(i - r * sorb);
    const u64 w = comb[r * L + (k >> 6)];
    states[i] = ((w >> (k & 63)) & 1ull) ? 1.0 : -1.0;
  }
}

template <int L, typename T, bool WITH_H>
static int launch_enumerate_L(const u64 *bra, const T *h1e, const T *h2e, u64 *comb, T *hmat, long long n, const ExcGeom &g,
                              cudaStream_t st) {
  const long long M = (long long)g.nsd + 1;
  const int tiles = (int)((M + 1 + kEnumTileRows - 1) / kEnumTileRows);
  const long long blocks = n * tiles;
  if (blocks > 0x7fffffffLL) {
    set_error("enumerate: n * tiles = %lld exceeds the grid limit; split the batch", blocks);
    return 1;
  }
  enumerate_kernel<L, T, WITH_H><<<(unsigned)blocks, kEnumThreads, 0, st>>>(bra, h1e, h2e, comb, hmat, n, tiles, g);
  count_launch();
  if (int rc = check_launch("enumerate_kernel")) return rc;
  if (WITH_H) {
    diag_kernel<L, T><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(bra, h1e, h2e, hmat, n, M, g.sorb, g.nele);
    count_launch();
    if (int rc = check_launch("diag_kernel")) return rc;
  }
  return 0;
}

template <typename T, bool WITH_H>
static int launch_enumerate_T(const u64 *bra, const T *h1e, const T *h2e, u64 *comb, T *hmat, long long n, const ExcGeom &g,
                              cudaStream_t st) {
  switch (g.L) {
    case 1: return launch_enumerate_L<1, T, WITH_H>(bra, h1e, h2e, comb, hmat, n, g, st);
    case 2: return launch_enumerate_L<2, T, WITH_H>(bra, h1e, h2e, comb, hmat, n, g, st);
    case 3: return launch_enumerate_L<3, T, WITH_H>(bra, h1e, h2e, comb, hmat, n, g, st);
  }
  set_error("unsupported ONV length L=%d", g.L);
  return 1;
}

int launch_comb(const u64 *bra, u64 *comb, long long n, const ExcGeom &g, cudaStream_t st) {
  return launch_enumerate_T<double, false>(bra, nullptr, nullptr, comb, nullptr, n, g, st);
}

int launch_comb_hij_f64(const u64 *bra, const double *h1e, const double *h2e, u64 *comb, double *hmat, long long n,
                        const ExcGeom &g, cudaStream_t st) {
  return launch_enumerate_T<double, true>(bra, h1e, h2e, comb, hmat, n, g, st);
}

int launch_comb_hij_f32(const u64 *bra, const float *h1e, const float *h2e, u64 *comb, float *hmat, long long n,
                        const ExcGeom &g, cudaStream_t st) {
  return launch_enumerate_T<float, true>(bra, h1e, h2e, comb, hmat, n, g, st);
}

// stand-alone diagonal (used by the fused local-energy op): out[s * stride] = <x_s|H|x_s>
int launch_diag_f64(const u64 *bra, const double *h1e, const double *h2e, double *out, long long n, long long stride, int L,
                    int sorb, int nele, cudaStream_t st) {
  if (n == 0) return 0;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  switch (L) {
    case 1: diag_kernel<1, double><<<blocks, 128, 0, st>>>(bra, h1e, h2e, out, n, stride, sorb, nele); break;
    case 2: diag_kernel<2, double><<<blocks, 128, 0, st>>>(bra, h1e, h2e, out, n, stride, sorb, nele); break;
    case 3: diag_kernel<3, double><<<blocks, 128, 0, st>>>(bra, h1e, h2e, out, n, stride, sorb, nele); break;
    default: set_error("unsupported ONV length L=%d", L); return 1;
  }
  count_launch();
  return check_launch("diag_kernel");
}

int launch_states(const u64 *comb, double *states, long long rows, int sorb, cudaStream_t st) {
  const int L = (sorb - 1) / 64 + 1;
  const long long total = rows * sorb;
  const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (total == 0) return 0;
  switch (L) {
    case 1: states_kernel<1><<<blocks, 256, 0, st>>>(comb, states, rows, sorb); break;
    case 2: states_kernel<2><<<blocks, 256, 0, st>>>(comb, states, rows, sorb); break;
    default: states_kernel<3><<<blocks, 256, 0, st>>>(comb, states, rows, sorb); break;
  }
  count_launch();
  return check_launch("states_kernel");
}

}  // namespace pynqs
