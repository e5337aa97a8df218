// table.cu -- the two steps either side of the local-energy kernels:
//   * sort of the unique-sample table: the reference orders the keys with 8L stable argsorts, one
//     per byte column (torch_lexsort, utils/public_function.py:626-689), which is the ascending
//     order of the little-endian multi-word integer with ties in input order.  Here: L stable LSD
//     radix passes over 64-bit words (cub onesweep, restricted to the significant bits of each
//     word), carrying a 32-bit permutation, then ONE gather of keys / psi / permutation.
//   * weighted moments of the local energy: the three reductions behind dist_mean / dist_var
//     (utils/stats/dist_stats.py:18-79) in one deterministic kernel.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace pynqs {

struct SortLayout {
  long long words[2], perm[2], cub, total;
  size_t cub_bytes;
};

static SortLayout sort_layout(long long N) {
  SortLayout l;
  size_t tmp = 0;
  cub::DoubleBuffer<u64> k(nullptr, nullptr);
  cub::DoubleBuffer<u32> v(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, k, v, (int)(N > 0 ? N : 1), 0, 64);
  const long long n8 = ((N + 1) & ~1LL);
  l.words[0] = 0;
  l.words[1] = l.words[0] + n8 * 8;
  l.perm[0] = l.words[1] + n8 * 8;
  l.perm[1] = l.perm[0] + n8 * 4;
  l.cub = l.perm[1] + n8 * 4;
  l.cub_bytes = tmp;
  l.total = l.cub + (long long)((tmp + 255) & ~(size_t)255) + 256;
  return l;
}

long long sort_workspace_bytes(long long N) { return sort_layout(N).total; }

// words[i] = word k of row perm[i] (perm == nullptr: identity, and the permutation is initialised)
__global__ void __launch_bounds__(256)
sort_gather_word_kernel(const u64 *__restrict__ key, long long N, int L, int k, const u32 *__restrict__ perm, u64 *__restrict__ words,
                        u32 *__restrict__ iota) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (perm == nullptr) {
    words[i] = key[i * L + k];
    iota[i] = (u32)i;
  } else {
    words[i] = key[(long long)perm[i] * L + k];
  }
}

template <int L, int PW>  // PW = psi width in 8-byte words (1 real, 2 complex)
__global__ void __launch_bounds__(256)
sort_apply_kernel(const u64 *__restrict__ key, const u64 *__restrict__ psi, long long N, const u32 *__restrict__ perm,
                  u64 *__restrict__ key_out, u64 *__restrict__ psi_out, long long *__restrict__ perm_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const long long s = perm[i];
#pragma unroll
  for (int w = 0; w < L; ++w) key_out[i * L + w] = key[s * L + w];
  if (psi != nullptr) {
#pragma unroll
    for (int w = 0; w < PW; ++w) psi_out[i * PW + w] = psi[s * PW + w];
  }
  if (perm_out != nullptr) perm_out[i] = s;
}

int launch_sort_table(const u64 *key, const void *psi, long long N, int L, int sorb, int psi_bytes, u64 *key_out, void *psi_out,
                      long long *perm_out, void *ws, long long ws_bytes, cudaStream_t st) {
  if (N >= (1LL << 31)) {
    set_error("table sort supports fewer than 2^31 keys (got %lld)", N);
    return 1;
  }
  if (N == 0) return 0;
  const SortLayout l = sort_layout(N);
  if (ws_bytes < l.total) {
    set_error("sort workspace too small: %lld < %lld bytes", ws_bytes, l.total);
    return 4;
  }
  char *b = static_cast<char *>(ws);
  cub::DoubleBuffer<u64> words(reinterpret_cast<u64 *>(b + l.words[0]), reinterpret_cast<u64 *>(b + l.words[1]));
  cub::DoubleBuffer<u32> perm(reinterpret_cast<u32 *>(b + l.perm[0]), reinterpret_cast<u32 *>(b + l.perm[1]));
  const unsigned blocks = (unsigned)((N + 255) / 256);
  for (int k = 0; k < L; ++k) {  // least-significant word first; every pass is stable
    sort_gather_word_kernel<<<blocks, 256, 0, st>>>(key, N, L, k, k == 0 ? nullptr : perm.Current(), words.Current(), perm.Current());
    count_launch();
    int bits = sorb - 64 * k;  // sorb <= 0: no promise about unused bits, sort on all 64
    if (sorb <= 0 || bits > 64) bits = 64;
    if (bits < 1) bits = 1;
    size_t tmp = l.cub_bytes;
    const cudaError_t e = cub::DeviceRadixSort::SortPairs(b + l.cub, tmp, words, perm, (int)N, 0, bits, st);
    if (e != cudaSuccess) {
      set_error("radix sort: CUDA error %d (%s)", (int)e, cudaGetErrorString(e));
      return 3;
    }
    count_launch((bits + 7) / 8 + 1);  // histogram + one onesweep pass per 8 bits
  }
  const u64 *p = static_cast<const u64 *>(psi);
  u64 *po = static_cast<u64 *>(psi_out);
#define PYNQS_APPLY(LL)                                                                                                   \
  if (psi_bytes == 16) sort_apply_kernel<LL, 2><<<blocks, 256, 0, st>>>(key, p, N, perm.Current(), key_out, po, perm_out); \
  else sort_apply_kernel<LL, 1><<<blocks, 256, 0, st>>>(key, p, N, perm.Current(), key_out, po, perm_out)
  switch (L) {
    case 1: PYNQS_APPLY(1); break;
    case 2: PYNQS_APPLY(2); break;
    case 3: PYNQS_APPLY(3); break;
    default: set_error("unsupported ONV length L=%d", L); return 1;
  }
#undef PYNQS_APPLY
  count_launch();
  return check_launch("table sort");
}

// ---- merge of per-rank sample counts --------------------------------------------------------------------
// merge_counts[idx[i]] += counts[i]  (merge_sample_cpu, cpp_src/tensor/cpu_tensor.cpp:537-556).  64-bit integer
// atomics: exact and order-independent (the reference's CUDA kernel adds without atomics, cuda/kernel.cu:520-536,
// and relies on split_idx to keep equal indices apart).  Indices outside [0, length) are ignored.
__global__ void __launch_bounds__(256)
merge_counts_kernel(const long long *__restrict__ idx, const long long *__restrict__ counts, long long n, long long length,
                    unsigned long long *__restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long k = idx[i];
    if (k >= 0 && k < length) atomicAdd(out + k, (unsigned long long)counts[i]);
  }
}

int launch_merge_counts(const long long *idx, const long long *counts, long long n, long long length, long long *out, cudaStream_t st) {
  if (length > 0 && cudaMemsetAsync(out, 0, 8 * (size_t)length, st) != cudaSuccess) return check_launch("merge memset");
  if (n == 0 || length == 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  merge_counts_kernel<<<(unsigned)blocks, 256, 0, st>>>(idx, counts, n, length, reinterpret_cast<unsigned long long *>(out));
  count_launch();
  return check_launch("merge_counts_kernel");
}

// ---- weighted moments ---------------------------------------------------------------------------
// out[0..6] = { sum w, sum w d_re, sum w d_im, sum w |d|^2, c_re, c_im, n }, d = E - c, c = E[0]
// (shifted moments: |d| is of the size of the spread, so the variance has no cancellation).
// w_i = weight_i (kind 0), weight_i^2 (kind 1, real amplitude) or |weight_i|^2 (kind 2, complex).
// Deterministic: fixed grid, fixed strides, block partials combined in block order by the last block.
constexpr int kMomBlocks = 296, kMomThreads = 256;

long long moments_scratch_bytes() { return (long long)(kMomBlocks * 4 + 2) * 8; }

template <bool ECPLX, int WKIND>
__global__ void __launch_bounds__(kMomThreads)
moments_kernel(const double *__restrict__ eloc, const double *__restrict__ weight, long long n, double *__restrict__ partial,
               unsigned int *__restrict__ ticket, double *__restrict__ out) {
  const double c_re = n > 0 ? eloc[0] : 0.0;
  const double c_im = (ECPLX && n > 0) ? eloc[1] : 0.0;
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (long long i = (long long)blockIdx.x * kMomThreads + threadIdx.x; i < n; i += (long long)kMomBlocks * kMomThreads) {
    double w;
    if (WKIND == 0) w = weight[i];
    else if (WKIND == 1) w = weight[i] * weight[i];
    else w = weight[2 * i] * weight[2 * i] + weight[2 * i + 1] * weight[2 * i + 1];
    const double dr = (ECPLX ? eloc[2 * i] : eloc[i]) - c_re;
    const double di = ECPLX ? eloc[2 * i + 1] - c_im : 0.0;
    s[0] += w;
    s[1] += w * dr;
    s[2] += w * di;
    s[3] += w * (dr * dr + di * di);
  }
  __shared__ double sm[kMomThreads / 32][4];
  __shared__ bool last;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) sm[threadIdx.x >> 5][k] = s[k];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < kMomThreads / 32; ++w) t += sm[w][threadIdx.x];
    partial[blockIdx.x * 4 + threadIdx.x] = t;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < 128) {  // one warp per component, fixed assignment and a fixed shuffle tree
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double t = 0.0;
    for (int bq = lane; bq < (int)gridDim.x; bq += 32) t += __ldcg(partial + bq * 4 + k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) out[k] = t;
  }
  if (threadIdx.x == 0) {
    out[4] = c_re;
    out[5] = c_im;
    out[6] = (double)n;
    *ticket = 0;  // ready for the next launch on this stream
  }
}

int launch_moments(const double *eloc, int eloc_complex, const double *weight, int weight_kind, long long n, void *scratch,
                   double *out, cudaStream_t st) {
  double *partial = static_cast<double *>(scratch);
  unsigned int *ticket = reinterpret_cast<unsigned int *>(partial + kMomBlocks * 4);
#define PYNQS_MOM(EC, WK) moments_kernel<EC, WK><<<kMomBlocks, kMomThreads, 0, st>>>(eloc, weight, n, partial, ticket, out)
  if (eloc_complex) {
    if (weight_kind == 0) PYNQS_MOM(true, 0);
    else if (weight_kind == 1) PYNQS_MOM(true, 1);
    else PYNQS_MOM(true, 2);
  } else {
    if (weight_kind == 0) PYNQS_MOM(false, 0);
    else if (weight_kind == 1) PYNQS_MOM(false, 1);
    else PYNQS_MOM(false, 2);
  }
#undef PYNQS_MOM
  count_launch();
  return check_launch("moments_kernel");
}

}  // namespace pynqs
