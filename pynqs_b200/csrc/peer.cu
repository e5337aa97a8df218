// peer.cu -- collectives of the multi-GPU step as pull kernels over NVLink peer memory.
//
// Every rank owns a symmetric buffer (torch.distributed._symmetric_memory: the same allocation mapped into every peer's
// address space; NVSwitch gives each GPU full bandwidth to each peer).  A rank publishes its piece by writing it into its
// own buffer; after a barrier every rank PULLS what it needs straight out of the peers' buffers:
//   peer_gather_kernel       all-gather: W pieces of equal size -> one contiguous tensor (the sample exchange: ONVs, psi);
//   peer_gather_rows_kernel  gather fused with the exchange of the local energies: rank r computed the energies of the
//                            r-th slice of the beta-grouped table; a rank needs the energies of ITS rows of the sorted
//                            table, which sit at scattered positions of that order -- it reads exactly those elements from
//                            whichever peer holds them (1/W of the data moves, and no all-gathered copy is ever stored).
// Replaces NCCL all_gather calls whose cost at this size (16 MB / 8 MB per step) is launch + protocol latency, not bandwidth.
#include "common.cuh"

namespace pynqs {

__global__ void __launch_bounds__(256)
peer_gather_kernel(const uint4 *const *__restrict__ peers, int world, long long src_off16, long long n16, uint4 *__restrict__ out) {
  const long long total = (long long)world * n16;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // four independent 16-byte loads in flight per thread: NVLink latency is ~1-2 us
  for (; i + 3 * stride < total; i += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long k = i + u * stride;
      const int r = (int)(k / n16);
      v[u] = peers[r][src_off16 + (k - (long long)r * n16)];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) out[i + u * stride] = v[u];
  }
  for (; i < total; i += stride) {
    const int r = (int)(i / n16);
    out[i] = peers[r][src_off16 + (i - (long long)r * n16)];
  }
}

// out[i] = element pos[i] of the concatenation of the peers' pieces; piece k holds q + (k < rem) elements
template <int WORDS>  // element size in 8-byte words
__global__ void __launch_bounds__(256)
peer_gather_rows_kernel(const u64 *const *__restrict__ peers, long long src_off8, const u32 *__restrict__ pos, long long n, long long q,
                        int rem, u64 *__restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long p = pos[i];
  // owner of global position p under split_length_idx: the first `rem` pieces are one element longer
  const long long big = (long long)rem * (q + 1);
  int r;
  long long local;
  if (p < big) {
    r = (int)(p / (q + 1));
    local = p - (long long)r * (q + 1);
  } else {
    r = rem + (int)((p - big) / q);
    local = p - big - (long long)(r - rem) * q;
  }
  const u64 *src = peers[r] + src_off8 + local * WORDS;
#pragma unroll
  for (int w = 0; w < WORDS; ++w) out[i * WORDS + w] = src[w];
}

int launch_peer_gather(const void *const *peers, int world, long long src_off, long long bytes, void *out, cudaStream_t st) {
  if (bytes == 0) return 0;
  if ((bytes & 15) || (src_off & 15)) {
    set_error("peer_gather: offsets and sizes must be multiples of 16 bytes (got %lld, %lld)", src_off, bytes);
    return 1;
  }
  const long long n16 = bytes / 16, total = n16 * world;
  long long blocks = (total + 256 * 4 - 1) / (256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  peer_gather_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const uint4 *const *>(peers), world, src_off / 16, n16,
                                                       static_cast<uint4 *>(out));
  count_launch();
  return check_launch("peer_gather_kernel");
}

int launch_peer_gather_rows(const void *const *peers, long long src_off, const u32 *pos, long long n, long long total, int world,
                            int elem_bytes, void *out, cudaStream_t st) {
  if (n == 0) return 0;
  if ((elem_bytes != 8 && elem_bytes != 16) || (src_off & 7)) {
    set_error("peer_gather_rows: element size %d / offset %lld unsupported", elem_bytes, src_off);
    return 1;
  }
  const long long q = total / world;
  const int rem = (int)(total - q * world);
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (elem_bytes == 16)
    peer_gather_rows_kernel<2><<<blocks, 256, 0, st>>>(reinterpret_cast<const u64 *const *>(peers), src_off / 8, pos, n, q, rem,
                                                    static_cast<u64 *>(out));
  else
    peer_gather_rows_kernel<1><<<blocks, 256, 0, st>>>(reinterpret_cast<const u64 *const *>(peers), src_off / 8, pos, n, q, rem,
                                                    static_cast<u64 *>(out));
  count_launch();
  return check_launch("peer_gather_rows_kernel");
}

}  // namespace pynqs
