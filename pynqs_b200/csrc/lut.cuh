// lut.cuh -- lookup of ONVs in the sorted unique-sample table (device side).
//
// classic_search: the reference's probe sequence (binary_search_BigInteger,
// cpp_src/tensor/cpu_tensor.cpp:589-640; BigInteger_device, cuda/kernel.cu:625-650).
// hashed_search: same answer for a sorted table without duplicates, one 16-byte bucket probe
// per query instead of ~log2(N) dependent loads.
#pragma once
#include "common.cuh"

namespace pynqs {

// workspace layout: HashHeader (256 B) | buckets[nb] of 32 B {tag[4], idx[4]}
struct HashHeader {
  u32 log2_nb;
  u32 has_dup;  // set by the build when two adjacent sorted keys are equal
  u64 n_keys;
  u32 pad[60];
};
static_assert(sizeof(HashHeader) == 256, "header is 256 bytes");

struct HashBucket {
  u32 tag[4];
  u32 idx[4];
};

template <int L>
__device__ __forceinline__ u64 hash_onv(const Onv<L> &x) {
  u64 h = x.w[0] * 0x9E3779B97F4A7C15ull;
#pragma unroll
  for (int i = 1; i < L; ++i) h = (h ^ (h >> 32) ^ x.w[i]) * 0xD6E8FEB86659FD93ull;
  h ^= h >> 32;
  h *= 0xD6E8FEB86659FD93ull;
  h ^= h >> 32;
  return h;
}

__device__ __forceinline__ u32 hash_tag(u64 h) {
  const u32 t = (u32)h;
  return t ? t : 1u;
}

template <int L>
__device__ __forceinline__ long long classic_search(const u64 *__restrict__ key, long long N, const Onv<L> &q) {
  long long lo = 0, hi = N - 1;
  while (lo <= hi) {
    const long long mid = lo + (hi - lo) / 2;
    const Onv<L> e = load_onv<L>(key + mid * L);
    const int c = cmp_onv<L>(e, q);
    if (c == 0) return mid;
    if (c < 0) lo = mid + 1;
    else hi = mid - 1;
  }
  return -1;
}

template <int L>
__device__ __forceinline__ long long hashed_search(const u64 *__restrict__ key, long long N, const HashHeader *__restrict__ hdr,
                                                   const Onv<L> &q) {
  if (hdr->has_dup) return classic_search<L>(key, N, q);
  const u32 log2_nb = hdr->log2_nb;
  const HashBucket *__restrict__ buckets = reinterpret_cast<const HashBucket *>(hdr + 1);
  const u64 h = hash_onv<L>(q);
  const u32 tag = hash_tag(h);
  const u32 mask = (1u << log2_nb) - 1u;
  u32 b = (u32)(h >> (64 - log2_nb));
  for (u32 probe = 0; probe <= mask; ++probe) {
    const uint4 t = __ldg(reinterpret_cast<const uint4 *>(buckets[b].tag));
    const u32 tg[4] = {t.x, t.y, t.z, t.w};
    bool open = false;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (tg[s] == tag) {
        const u32 id = __ldg(&buckets[b].idx[s]);
        const Onv<L> e = load_onv<L>(key + (long long)id * L);
        if (eq_onv<L>(e, q)) return (long long)id;
      }
      open |= (tg[s] == 0u);
    }
    if (open) return -1;
    b = (b + 1) & mask;
  }
  return -1;
}

}  // namespace pynqs
