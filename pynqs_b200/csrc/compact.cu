// compact.cu -- the index glue either side of the lookup, as kernels:
//
//   * WavefunctionLUT.lookup (utils/public_function.py:817-838) follows wavefunction_lut with
//       baseline = arange(n); onv_idx = baseline[mask]; onv_not_idx = baseline[~mask]; value = wf_value[idx[mask]]
//     -- an arange, two boolean-mask selections (each a flagged select over n) and a gather.  Here: one pass that counts
//     the hits per block of 2048 queries, a scan of the block counts, and one pass that writes the three outputs in order.
//   * Func (vmc/energy/flip.py:44-61) removes duplicate rows among the LUT misses with torch.unique(dim=0,
//     return_inverse=True) before calling the ansatz.  Here: the library's radix sort of the rows (table.cu), then head flags
//     -> block counts -> scan -> one pass that writes the distinct rows and, through the sort permutation, the inverse map.
//     The distinct rows come out in ascending ONV order (torch.unique orders rows byte 0 first); Func only needs
//     unique[inverse] == x, which holds for either order.
//
// Deterministic: ranks come from ballots and fixed-order prefixes, never from atomics.
#include "common.cuh"

namespace pynqs {

constexpr int kCpThreads = 256, kCpItems = 8, kCpBlock = kCpThreads * kCpItems;

// blk[b] = number of set flags among items [b * kCpBlock, (b + 1) * kCpBlock)
__global__ void __launch_bounds__(kCpThreads)
count_flags_kernel(const unsigned char *__restrict__ flag, long long n, u32 *__restrict__ blk) {
  __shared__ u32 wsum[kCpThreads / 32];
  const long long base = (long long)blockIdx.x * kCpBlock;
  u32 c = 0;
#pragma unroll
  for (int j = 0; j < kCpItems; ++j) {
    const long long i = base + j * kCpThreads + threadIdx.x;
    c += (i < n && flag[i]) ? 1u : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    u32 t = 0;
    for (int w = 0; w < kCpThreads / 32; ++w) t += wsum[w];
    blk[blockIdx.x] = t;
  }
}

// off[b] = sum of blk[0 .. b) (64-bit), total[0] = sum of all: one CTA walks the block counts 1024 at a time
__global__ void __launch_bounds__(1024) scan_blocks_kernel(const u32 *__restrict__ blk, long long nblk, unsigned long long *__restrict__ off,
                                                           unsigned long long *__restrict__ total) {
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned long long carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0ull;
  __syncthreads();
  for (long long b0 = 0; b0 < nblk; b0 += 1024) {
    const long long b = b0 + threadIdx.x;
    const unsigned long long v = b < nblk ? (unsigned long long)blk[b] : 0ull;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned long long before = carry;
    for (int w = 0; w < warp; ++w) before += wsum[w];
    if (b < nblk) off[b] = before + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) total[0] = carry;
}

// rank of this thread's item among the set flags of its block, items taken in index order: item j of thread t is element
// j * kCpThreads + t, so the order is j-major; returns the exclusive rank for each of the thread's items
__device__ __forceinline__ void block_ranks(const bool (&f)[kCpItems], u32 (&rank)[kCpItems], u32 *wsum /* [kCpItems][warps] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kCpThreads / 32;
  u32 bal[kCpItems];
#pragma unroll
  for (int j = 0; j < kCpItems; ++j) {
    bal[j] = __ballot_sync(0xffffffffu, f[j]);
    if (lane == 0) wsum[j * kWarps + warp] = (u32)__popc(bal[j]);
  }
  __syncthreads();
  u32 before = 0;
#pragma unroll
  for (int j = 0; j < kCpItems; ++j) {
    u32 mine = before;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const u32 t = wsum[j * kWarps + w];
      if (w < warp) mine += t;
      before += t;
    }
    rank[j] = mine + (u32)__popc(bal[j] & ((1u << lane) - 1u));
  }
  __syncthreads();
}

template <int PW>  // value width in 8-byte words
__global__ void __launch_bounds__(kCpThreads)
lookup_emit_kernel(const unsigned char *__restrict__ mask, const long long *__restrict__ idx, long long n,
                   const unsigned long long *__restrict__ off, const u64 *__restrict__ psi, long long *__restrict__ hit_pos,
                   long long *__restrict__ miss_pos, u64 *__restrict__ value) {
  __shared__ u32 wsum[kCpItems * (kCpThreads / 32)];
  const long long base = (long long)blockIdx.x * kCpBlock;
  bool f[kCpItems];
  u32 rank[kCpItems];
#pragma unroll
  for (int j = 0; j < kCpItems; ++j) {
    const long long i = base + j * kCpThreads + threadIdx.x;
    f[j] = i < n && mask[i];
  }
  block_ranks(f, rank, wsum);
  const long long h0 = (long long)off[blockIdx.x], m0 = base - h0;
#pragma unroll
  for (int j = 0; j < kCpItems; ++j) {
    const long long local = j * kCpThreads + threadIdx.x, i = base + local;
    if (i >= n) continue;
    if (f[j]) {
      const long long o = h0 + rank[j], row = idx[i];
      hit_pos[o] = i;
#pragma unroll
      for (int w = 0; w < PW; ++w) value[o * PW + w] = psi[row * PW + w];
    } else {
      miss_pos[m0 + (local - rank[j])] = i;
    }
  }
}

template <int L>
__global__ void __launch_bounds__(256)
head_flags_kernel(const u64 *__restrict__ key, long long n, unsigned char *__restrict__ flag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = i == 0 || !eq_onv<L>(load_onv<L>(key + i * L), load_onv<L>(key + (i - 1) * L));
}

template <int L>
__global__ void __launch_bounds__(kCpThreads)
unique_emit_kernel(const u64 *__restrict__ key, const long long *__restrict__ perm, const unsigned char *__restrict__ flag, long long n,
                   const unsigned long long *__restrict__ off, u64 *__restrict__ uniq, long long *__restrict__ inverse) {
  __shared__ u32 wsum[kCpItems * (kCpThreads / 32)];
  const long long base = (long long)blockIdx.x * kCpBlock;
  bool f[kCpItems];
  u32 rank[kCpItems];
#pragma unroll
  for (int j = 0; j < kCpItems; ++j) {
    const long long i = base + j * kCpThreads + threadIdx.x;
    f[j] = i < n && flag[i];
  }
  block_ranks(f, rank, wsum);
  const long long g0 = (long long)off[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kCpItems; ++j) {
    const long long i = base + j * kCpThreads + threadIdx.x;
    if (i >= n) continue;
    // group of row i = number of heads at or before i, minus one
    const long long gid = g0 + rank[j] + (f[j] ? 1 : 0) - 1;
    if (f[j]) {
#pragma unroll
      for (int w = 0; w < L; ++w) uniq[gid * L + w] = key[i * L + w];
    }
    inverse[perm[i]] = gid;
  }
}

// ---- host side --------------------------------------------------------------------------------------------------------
struct CompactLayout {
  long long blk, off, total_cnt, flag, bytes;
  long long nblk;
};
static CompactLayout compact_layout(long long n, bool with_flags) {
  CompactLayout l;
  l.nblk = (n + kCpBlock - 1) / kCpBlock;
  auto up = [](long long v) { return (v + 255) & ~255LL; };
  l.blk = 0;
  l.off = up(4 * l.nblk);
  l.total_cnt = l.off + up(8 * l.nblk);
  l.flag = l.total_cnt + 256;
  l.bytes = l.flag + (with_flags ? up(n) : 0) + 256;
  return l;
}

long long compact_scratch_bytes(long long n) { return compact_layout(n < 0 ? 0 : n, true).bytes; }

// count: *total (device, 8 bytes inside the scratch; also copied to total_out if non-null) = number of set flags
static int count_and_scan(const unsigned char *flag, long long n, char *sc, const CompactLayout &l, unsigned long long *total_out,
                          cudaStream_t st) {
  u32 *blk = reinterpret_cast<u32 *>(sc + l.blk);
  unsigned long long *off = reinterpret_cast<unsigned long long *>(sc + l.off);
  unsigned long long *tot = reinterpret_cast<unsigned long long *>(sc + l.total_cnt);
  if (l.nblk > 0x7fffffffLL) {
    set_error("compaction: too many rows (%lld)", n);
    return 1;
  }
  count_flags_kernel<<<(unsigned)l.nblk, kCpThreads, 0, st>>>(flag, n, blk);
  scan_blocks_kernel<<<1, 1024, 0, st>>>(blk, l.nblk, off, tot);
  count_launch(2);
  if (int rc = check_launch("compaction count / scan")) return rc;
  if (total_out != nullptr && cudaMemcpyAsync(total_out, tot, 8, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
    return check_launch("compaction total copy");
  return 0;
}

int launch_lookup_count(const unsigned char *mask, long long n, void *scratch, long long scratch_bytes, unsigned long long *n_hit, cudaStream_t st) {
  if (n == 0) return cudaMemsetAsync(n_hit, 0, 8, st) == cudaSuccess ? 0 : check_launch("lookup_count memset");
  const CompactLayout l = compact_layout(n, true);
  if (scratch_bytes < l.bytes) {
    set_error("lookup compaction scratch too small: %lld < %lld bytes", scratch_bytes, l.bytes);
    return 4;
  }
  return count_and_scan(mask, n, static_cast<char *>(scratch), l, n_hit, st);
}

int launch_lookup_emit(const unsigned char *mask, const long long *idx, long long n, const void *psi, int psi_bytes, void *scratch,
                       long long *hit_pos, long long *miss_pos, void *value, cudaStream_t st) {
  if (n == 0) return 0;
  const CompactLayout l = compact_layout(n, true);
  const unsigned long long *off = reinterpret_cast<const unsigned long long *>(static_cast<char *>(scratch) + l.off);
  if (psi_bytes == 16)
    lookup_emit_kernel<2><<<(unsigned)l.nblk, kCpThreads, 0, st>>>(mask, idx, n, off, static_cast<const u64 *>(psi), hit_pos, miss_pos,
                                                                static_cast<u64 *>(value));
  else
    lookup_emit_kernel<1><<<(unsigned)l.nblk, kCpThreads, 0, st>>>(mask, idx, n, off, static_cast<const u64 *>(psi), hit_pos, miss_pos,
                                                                static_cast<u64 *>(value));
  count_launch();
  return check_launch("lookup_emit_kernel");
}

int launch_unique_count(const u64 *sorted_key, long long n, int L, void *scratch, long long scratch_bytes, unsigned long long *n_unique,
                        cudaStream_t st) {
  if (n == 0) return cudaMemsetAsync(n_unique, 0, 8, st) == cudaSuccess ? 0 : check_launch("unique_count memset");
  const CompactLayout l = compact_layout(n, true);
  if (scratch_bytes < l.bytes) {
    set_error("unique scratch too small: %lld < %lld bytes", scratch_bytes, l.bytes);
    return 4;
  }
  char *sc = static_cast<char *>(scratch);
  unsigned char *flag = reinterpret_cast<unsigned char *>(sc + l.flag);
  const unsigned blocks = (unsigned)((n + 255) / 256);
  switch (L) {
    case 1: head_flags_kernel<1><<<blocks, 256, 0, st>>>(sorted_key, n, flag); break;
    case 2: head_flags_kernel<2><<<blocks, 256, 0, st>>>(sorted_key, n, flag); break;
    case 3: head_flags_kernel<3><<<blocks, 256, 0, st>>>(sorted_key, n, flag); break;
    default: set_error("unsupported ONV length L=%d", L); return 1;
  }
  count_launch();
  return count_and_scan(flag, n, sc, l, n_unique, st);
}

int launch_unique_emit(const u64 *sorted_key, const long long *perm, long long n, int L, void *scratch, u64 *uniq, long long *inverse,
                       cudaStream_t st) {
  if (n == 0) return 0;
  const CompactLayout l = compact_layout(n, true);
  char *sc = static_cast<char *>(scratch);
  const unsigned char *flag = reinterpret_cast<const unsigned char *>(sc + l.flag);
  const unsigned long long *off = reinterpret_cast<const unsigned long long *>(sc + l.off);
  switch (L) {
    case 1: unique_emit_kernel<1><<<(unsigned)l.nblk, kCpThreads, 0, st>>>(sorted_key, perm, flag, n, off, uniq, inverse); break;
    case 2: unique_emit_kernel<2><<<(unsigned)l.nblk, kCpThreads, 0, st>>>(sorted_key, perm, flag, n, off, uniq, inverse); break;
    case 3: unique_emit_kernel<3><<<(unsigned)l.nblk, kCpThreads, 0, st>>>(sorted_key, perm, flag, n, off, uniq, inverse); break;
    default: set_error("unsupported ONV length L=%d", L); return 1;
  }
  count_launch();
  return check_launch("unique_emit_kernel");
}

}  // namespace pynqs
