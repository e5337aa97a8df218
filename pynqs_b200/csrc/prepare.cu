// prepare.cu -- gather-friendly internal copies of the packed integrals.
//
// The reference layout (cpp_src/cpu/hamiltonian.cpp:13-31) makes the 32 lanes of a warp, which
// enumerate consecutive excitations, hit 32 different 128-byte lines of the triangular h2e array.
// These kernels copy the SAME numbers (bit for bit, signs of two_body() included) into tables
// whose fastest index is the one that varies across lanes:
//   T_ab[hB][pB][pA][hA]   alpha-beta doubles      (ranks = orbital >> 1)
//   T_aa[pp][hp], T_bb     same-spin doubles       (pp/hp = triangular pair index of the ranks)
//   T_s [k][spin][p][h]    <hk||pk> with its sign  (single excitations; p, h = same-spin ranks, so
//                          the lanes of a warp -- consecutive holes h of 2-3 particles p -- share
//                          a few lines at every step k of the ordered sum)
// Built once per Hamiltonian (the Python shim caches it per h2e tensor); 2.4 MB at 40 spin
// orbitals, L2-resident.
#include "prepare.cuh"

namespace pynqs {

template <typename T>
__global__ void __launch_bounds__(256)
prepare_kernel(const T *__restrict__ h2e, T *__restrict__ t_ab, T *__restrict__ t_aa, T *__restrict__ t_bb, T *__restrict__ t_s,
               int sorb) {
  const long long na = sorb / 2;
  const long long npair = na * (na - 1) / 2;
  const long long n_ab = na * na * na * na, n_pp = npair * npair, n_s = (long long)sorb * 2 * na * na;
  const long long total = n_ab + 2 * n_pp + n_s;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    if (e < n_ab) {
      const u32 hA = (u32)(e % na), pA = (u32)((e / na) % na), pB = (u32)((e / (na * na)) % na), hB = (u32)(e / (na * na * na));
      const u32 ha = 2 * hA, pa = 2 * pA, hb = 2 * hB + 1, pb = 2 * pB + 1;
      const u32 hh = ha > hb ? ha : hb, hl = ha > hb ? hb : ha, ph = pa > pb ? pa : pb, pl = pa > pb ? pb : pa;
      t_ab[e] = __ldg(h2e + pair_offset(hh, hl, ph, pl));
    } else if (e < n_ab + 2 * n_pp) {
      long long f = e - n_ab;
      const bool beta = f >= n_pp;
      if (beta) f -= n_pp;
      const int hp = (int)(f % npair), pp = (int)(f / npair);
      int a, b, c, d;
      tri_unpack(hp, a, b);
      tri_unpack(pp, c, d);
      const u32 s = beta ? 1u : 0u;
      const T v = __ldg(h2e + pair_offset(2 * a + s, 2 * b + s, 2 * c + s, 2 * d + s));
      (beta ? t_bb : t_aa)[f] = v;
    } else {
      const long long f = e - n_ab - 2 * n_pp;
      const u32 hr = (u32)(f % na), pr = (u32)((f / na) % na), spin = (u32)((f / (na * na)) & 1), k = (u32)(f / (2 * na * na));
      t_s[f] = two_body<T>(h2e, 2 * hr + spin, k, 2 * pr + spin, k);
    }
  }
}

long long prepared_bytes(int sorb, int dtype) { return prep_layout(sorb).total * (dtype == 1 ? 8 : 4); }

int launch_prepare(const void *h2e, int sorb, int dtype, void *ws, long long ws_bytes, cudaStream_t st) {
  const long long need = prepared_bytes(sorb, dtype);
  if (ws_bytes < need) {
    set_error("prepared-integral workspace too small: %lld < %lld bytes", ws_bytes, need);
    return 4;
  }
  const PrepLayout p = prep_layout(sorb);
  long long want = (p.total + 255) / 256;
  const unsigned blocks = (unsigned)(want < 148LL * 32 ? (want < 1 ? 1 : want) : 148LL * 32);
  if (dtype == 1) {
    double *b = reinterpret_cast<double *>(ws);
    prepare_kernel<double><<<blocks, 256, 0, st>>>((const double *)h2e, b + p.off_ab, b + p.off_aa, b + p.off_bb, b + p.off_s, sorb);
  } else {
    float *b = reinterpret_cast<float *>(ws);
    prepare_kernel<float><<<blocks, 256, 0, st>>>((const float *)h2e, b + p.off_ab, b + p.off_aa, b + p.off_bb, b + p.off_s, sorb);
  }
  count_launch();
  return check_launch("prepare_kernel");
}

}  // namespace pynqs
