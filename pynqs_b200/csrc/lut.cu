// lut.cu -- wavefunction_lut kernels: classic binary search (K6 replacement, cuda/kernel.cu:608-680)
// and the hash index (build + probe) used to accelerate large query batches.
#include "lut.cuh"

namespace pynqs {

template <int L>
__global__ void __launch_bounds__(256)
lut_classic_kernel(const u64 *__restrict__ key, long long N, const u64 *__restrict__ q, long long n, long long *__restrict__ idx,
                   unsigned char *__restrict__ mask) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const Onv<L> x = load_onv<L>(q + t * L);
    const long long r = classic_search<L>(key, N, x);
    idx[t] = r;
    mask[t] = r >= 0;
  }
}

template <int L>
__global__ void __launch_bounds__(256)
lut_hashed_kernel(const u64 *__restrict__ key, long long N, const HashHeader *__restrict__ hdr, const u64 *__restrict__ q,
                  long long n, long long *__restrict__ idx, unsigned char *__restrict__ mask) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const Onv<L> x = load_onv<L>(q + t * L);
    const long long r = hashed_search<L>(key, N, hdr, x);
    idx[t] = r;
    mask[t] = r >= 0;
  }
}

__global__ void hash_header_kernel(HashHeader *hdr, u32 log2_nb, u64 N) {
  hdr->log2_nb = log2_nb;
  hdr->has_dup = 0;
  hdr->n_keys = N;
}

template <int L>
__global__ void __launch_bounds__(256)
hash_build_kernel(const u64 *__restrict__ key, long long N, HashHeader *hdr) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const Onv<L> x = load_onv<L>(key + i * L);
  if (i > 0) {
    const Onv<L> prev = load_onv<L>(key + (i - 1) * L);
    if (eq_onv<L>(prev, x)) atomicExch(&hdr->has_dup, 1u);
  }
  const u32 log2_nb = hdr->log2_nb;
  HashBucket *buckets = reinterpret_cast<HashBucket *>(hdr + 1);
  const u64 h = hash_onv<L>(x);
  const u32 tag = hash_tag(h);
  const u32 mask = (1u << log2_nb) - 1u;
  u32 b = (u32)(h >> (64 - log2_nb));
  for (u32 probe = 0; probe <= mask; ++probe) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (atomicCAS(&buckets[b].tag[s], 0u, tag) == 0u) {
        buckets[b].idx[s] = (u32)i;
        return;
      }
    }
    b = (b + 1) & mask;
  }
}

static inline unsigned grid_for(long long n, int threads, long long cap) {
  long long want = (n + threads - 1) / threads;
  if (want < 1) want = 1;
  return (unsigned)(want < cap ? want : cap);
}

static u32 hash_log2_buckets(long long N) {
  u32 lg = 6;
  while ((1LL << lg) < N && lg < 31) ++lg;
  return lg;
}

long long hash_workspace_bytes(long long N) { return (long long)sizeof(HashHeader) + ((long long)sizeof(HashBucket) << hash_log2_buckets(N)); }

int launch_lut_classic(const u64 *key, long long N, const u64 *q, long long n, int L, long long *idx, unsigned char *mask,
                       cudaStream_t st) {
  if (n == 0) return 0;
  const unsigned blocks = grid_for(n, 256, 148LL * 64);
  switch (L) {
    case 1: lut_classic_kernel<1><<<blocks, 256, 0, st>>>(key, N, q, n, idx, mask); break;
    case 2: lut_classic_kernel<2><<<blocks, 256, 0, st>>>(key, N, q, n, idx, mask); break;
    case 3: lut_classic_kernel<3><<<blocks, 256, 0, st>>>(key, N, q, n, idx, mask); break;
    default: set_error("unsupported ONV length L=%d", L); return 1;
  }
  count_launch();
  return check_launch("lut_classic_kernel");
}

int launch_hash_build(const u64 *key, long long N, int L, void *ws, long long ws_bytes, cudaStream_t st) {
  if (N >= (1LL << 32)) {
    set_error("hash index supports fewer than 2^32 keys (got %lld)", N);
    return 1;
  }
  const long long need = hash_workspace_bytes(N);
  if (ws_bytes < need) {
    set_error("hash workspace too small: %lld < %lld bytes", ws_bytes, need);
    return 4;
  }
  if (cudaMemsetAsync(ws, 0, (size_t)need, st) != cudaSuccess) return check_launch("hash memset");
  HashHeader *hdr = reinterpret_cast<HashHeader *>(ws);
  hash_header_kernel<<<1, 1, 0, st>>>(hdr, hash_log2_buckets(N), (u64)N);
  count_launch();
  if (N > 0) {
    const unsigned blocks = (unsigned)((N + 255) / 256);
    switch (L) {
      case 1: hash_build_kernel<1><<<blocks, 256, 0, st>>>(key, N, hdr); break;
      case 2: hash_build_kernel<2><<<blocks, 256, 0, st>>>(key, N, hdr); break;
      case 3: hash_build_kernel<3><<<blocks, 256, 0, st>>>(key, N, hdr); break;
      default: set_error("unsupported ONV length L=%d", L); return 1;
    }
    count_launch();
  }
  return check_launch("hash_build_kernel");
}

int launch_lut_hashed(const u64 *key, long long N, const u64 *q, long long n, int L, const void *ws, long long *idx,
                      unsigned char *mask, cudaStream_t st) {
  if (n == 0) return 0;
  const HashHeader *hdr = reinterpret_cast<const HashHeader *>(ws);
  const unsigned blocks = grid_for(n, 256, 148LL * 64);
  switch (L) {
    case 1: lut_hashed_kernel<1><<<blocks, 256, 0, st>>>(key, N, hdr, q, n, idx, mask); break;
    case 2: lut_hashed_kernel<2><<<blocks, 256, 0, st>>>(key, N, hdr, q, n, idx, mask); break;
    case 3: lut_hashed_kernel<3><<<blocks, 256, 0, st>>>(key, N, hdr, q, n, idx, mask); break;
    default: set_error("unsupported ONV length L=%d", L); return 1;
  }
  count_launch();
  return check_launch("lut_hashed_kernel");
}

}  // namespace pynqs
