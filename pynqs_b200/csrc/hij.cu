// hij.cu -- <bra|H|ket> between explicit determinant lists (get_hij_torch).
//
// Replaces K3/K4 (get_Hij_kernel_3D / _2D, cuda/kernel.cu:72-128): the excitation is re-derived
// from the two bit strings (diff_type onstate.cpp:10-20, diff_orb :34-55, dispatch
// hamiltonian.cpp:87-102).  The reference tiles 32x32 with the sample index fastest, which
// makes adjacent threads stride M*L words; here the ket index is fastest so a warp reads
// consecutive ket rows and writes consecutive outputs.
#include "rederive.cuh"

namespace pynqs {

template <int L, typename T>
__global__ void __launch_bounds__(256)
hij_kernel(const u64 *__restrict__ bra, const u64 *__restrict__ ket, const T *__restrict__ h1e, const T *__restrict__ h2e,
           T *__restrict__ out, long long n, long long m, int ket3d, int sorb, int nele) {
  const long long total = n * m;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / m, j = t - i * m;
    const Onv<L> x = load_onv<L>(bra + i * L);
    const Onv<L> y = load_onv<L>(ket + (ket3d ? t : j) * L);
    out[t] = rederived_element<L, T>(x, y, h1e, h2e, sorb, nele);
  }
}

// 2-D mode (CI / hybrid callers: ci_vmc/hybrid.py:194, utils/ci/wavefunction.py:92): out[i, j] = <bra_i|H|ket_j>.  A CTA
// takes 256 kets -- one per thread, loaded ONCE into registers -- and a tile of kHijBraTile bras staged in shared memory
// (read back as broadcasts), so every ket row is fetched n / kHijBraTile times instead of n times and a warp writes 32
// consecutive outputs of a row.  Most pairs differ by more than a double excitation and cost two popcounts.
constexpr int kHijBraTile = 64;

template <int L, typename T>
__global__ void __launch_bounds__(256)
hij_2d_tiled_kernel(const u64 *__restrict__ bra, const u64 *__restrict__ ket, const T *__restrict__ h1e, const T *__restrict__ h2e,
                    T *__restrict__ out, long long n, long long m, int sorb, int nele) {
  __shared__ u64 s_bra[kHijBraTile * L];
  const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long i0 = (long long)blockIdx.y * kHijBraTile;
  const int rows = (int)(n - i0 < kHijBraTile ? n - i0 : kHijBraTile);
  for (int t = threadIdx.x; t < rows * L; t += 256) s_bra[t] = bra[i0 * L + t];
  __syncthreads();
  if (j >= m) return;
  const Onv<L> y = load_onv<L>(ket + j * L);
  for (int r = 0; r < rows; ++r) {
    Onv<L> x;
#pragma unroll
    for (int w = 0; w < L; ++w) x.w[w] = s_bra[r * L + w];
    out[(i0 + r) * m + j] = rederived_element<L, T>(x, y, h1e, h2e, sorb, nele);
  }
}

template <typename T>
int launch_hij(const u64 *bra, const u64 *ket, const T *h1e, const T *h2e, T *out, long long n, long long m, int ket3d,
               int sorb, int nele, cudaStream_t st) {
  const int L = (sorb - 1) / 64 + 1;
  const long long total = n * m;
  if (total == 0) return 0;
  if (!ket3d && m >= 256) {
    const long long gx = (m + 255) / 256, gy = (n + kHijBraTile - 1) / kHijBraTile;
    if (gy <= 65535) {
      const dim3 grid((unsigned)gx, (unsigned)gy);
      switch (L) {
        case 1: hij_2d_tiled_kernel<1, T><<<grid, 256, 0, st>>>(bra, ket, h1e, h2e, out, n, m, sorb, nele); break;
        case 2: hij_2d_tiled_kernel<2, T><<<grid, 256, 0, st>>>(bra, ket, h1e, h2e, out, n, m, sorb, nele); break;
        case 3: hij_2d_tiled_kernel<3, T><<<grid, 256, 0, st>>>(bra, ket, h1e, h2e, out, n, m, sorb, nele); break;
        default: set_error("unsupported ONV length L=%d", L); return 1;
      }
      count_launch();
      return check_launch("hij_2d_tiled_kernel");
    }
  }
  long long want = (total + 255) / 256;
  const unsigned blocks = (unsigned)(want < 148LL * 64 ? want : 148LL * 64);
  switch (L) {
    case 1: hij_kernel<1, T><<<blocks, 256, 0, st>>>(bra, ket, h1e, h2e, out, n, m, ket3d, sorb, nele); break;
    case 2: hij_kernel<2, T><<<blocks, 256, 0, st>>>(bra, ket, h1e, h2e, out, n, m, ket3d, sorb, nele); break;
    case 3: hij_kernel<3, T><<<blocks, 256, 0, st>>>(bra, ket, h1e, h2e, out, n, m, ket3d, sorb, nele); break;
    default: set_error("unsupported ONV length L=%d", L); return 1;
  }
  count_launch();
  return check_launch("hij_kernel");
}

template int launch_hij<double>(const u64 *, const u64 *, const double *, const double *, double *, long long, long long,
                                int, int, int, cudaStream_t);
template int launch_hij<float>(const u64 *, const u64 *, const float *, const float *, float *, long long, long long, int,
                               int, int, cudaStream_t);

}  // namespace pynqs
