// lut.cuh -- lookup of ONVs in the sorted unique-sample table (device side).
//
// classic_search: the reference's probe sequence (binary_search_BigInteger,
// cpp_src/tensor/cpu_tensor.cpp:589-640; BigInteger_device, cuda/kernel.cu:625-650).
//
// String-grouped index: same answers for a sorted table without duplicates, built for the access
// pattern of the local-energy path.  An ONV is an (alpha string, beta string) pair (even / odd
// bits).  The connected determinants of one sample that a warp probes together share one of the
// two strings (alpha-beta doubles with the same beta excitation, alpha-alpha doubles and alpha
// singles share the beta string; beta-beta doubles and beta singles share the alpha string), so
// the index groups the keys twice:
//   directory[B]: hash(beta string)  -> region of the keys with that beta string, hashed by alpha string
//   directory[A]: hash(alpha string) -> region of the keys with that alpha string, hashed by beta string
// A region is a power-of-two run of buckets {tag[4]} + {idx[4]} (tag = low 32 bits of the
// other string's hash, idx = row in the sorted table).  The 32 probes of a warp land in ONE
// region (a few 128-byte lines) instead of 32 random lines of a flat table, and whole groups of
// rows are skipped when their string is absent from the directory.  Every hit is verified
// against the key table, so hash collisions (even of the 64-bit directory keys, which merely
// merge two groups) cannot change a result.
#pragma once
#include "common.cuh"

namespace pynqs {

constexpr u64 kDirEmpty = ~0ull;

struct HashHeader {
  u32 log2_dir;  // directory slots (each of the two directories)
  u32 has_dup;   // set by the build when two adjacent sorted keys are equal
  u64 n_keys;
  u32 cursor;    // bump allocator over the shared bucket pool
  u32 pool_buckets;
  u32 n_claimed;  // build: directory slots in use (both directories), the work list of the carve step
  u32 pad[57];
};
static_assert(sizeof(HashHeader) == 256, "header is 256 bytes");

struct __align__(16) DirSlot {
  u64 h;    // 64-bit hash of the grouping string (kDirEmpty = free)
  u32 off;  // first bucket of the region in the pool
  u32 lg;   // log2(buckets of the region); during the build: number of keys of the group
};

// tag = low 32 bits of the string hash with bits 0 and 1 forced to 1 (never 0 = empty).  Slots fill
// from 0 upwards.  When a key finds its bucket full and moves on, it clears bit 0 of tag[3]: that is
// the bucket's OVERFLOW flag, the only case in which a probe has to look at the next bucket.
// Tags and rows live in two parallel arrays (uint4 per bucket each): the fast path of a probe reads
// only the 16 tag bytes, so the buckets of a region are packed 8 per 128-byte line.
typedef uint4 TagBucket;  // tag[4]
typedef uint4 IdxBucket;  // idx[4]

// workspace: header | dir[B] | dir[A] | tag pool (4N + 2 buckets) | idx pool | build scratch (4N u32)
struct IndexLayout {
  u32 log2_dir;
  long long dir_off[2], pool_off, idx_off, scratch_off, total;
  long long pool_buckets;
};

__host__ __device__ inline IndexLayout index_layout(long long N) {
  IndexLayout l;
  u32 lg = 6;
  while ((1LL << lg) < 2 * N && lg < 31) ++lg;
  l.log2_dir = lg;
  l.dir_off[0] = (long long)sizeof(HashHeader);
  l.dir_off[1] = l.dir_off[0] + ((long long)sizeof(DirSlot) << lg);
  l.pool_off = l.dir_off[1] + ((long long)sizeof(DirSlot) << lg);
  l.pool_buckets = 4 * N + 2;  // each of the two groupings needs < 2N buckets (<= 1 key per bucket on average)
  l.idx_off = l.pool_off + l.pool_buckets * (long long)sizeof(TagBucket);
  l.scratch_off = l.idx_off + l.pool_buckets * (long long)sizeof(IdxBucket);
  l.total = l.scratch_off + 4 * N * 4 + 16;  // per-key slots (2N) + list of claimed slots (<= 2N)
  return l;
}

struct IndexView {
  const HashHeader *hdr;
  const DirSlot *dir[2];  // [0] grouped by beta string, [1] grouped by alpha string
  const TagBucket *pool;  // tags
  const IdxBucket *rows;  // table rows, same indexing
  u32 log2_dir;
};

__host__ __device__ inline IndexView index_view(const void *ws, long long N) {
  const IndexLayout l = index_layout(N);
  const char *b = reinterpret_cast<const char *>(ws);
  IndexView v;
  v.hdr = reinterpret_cast<const HashHeader *>(b);
  v.dir[0] = reinterpret_cast<const DirSlot *>(b + l.dir_off[0]);
  v.dir[1] = reinterpret_cast<const DirSlot *>(b + l.dir_off[1]);
  v.pool = reinterpret_cast<const TagBucket *>(b + l.pool_off);
  v.rows = reinterpret_cast<const IdxBucket *>(b + l.idx_off);
  v.log2_dir = l.log2_dir;
  return v;
}

// 64-bit hash of one spin string: the words of the ONV masked to the even (alpha) or odd (beta)
// bits.  Two independent 32-bit multiply-xorshift mixes of the word halves (32-bit integer ops only:
// a 64-bit multiply costs ~4 of them on the GPU): the LOW word feeds the bucket tags, the HIGH word
// the directory slot and the bucket index inside a region.  Never returns kDirEmpty.
template <int L>
__device__ __forceinline__ u64 hash_string(const Onv<L> &x, u64 spin_mask) {
  u32 a = (u32)(x.w[0] & spin_mask), b = (u32)((x.w[0] & spin_mask) >> 32);
#pragma unroll
  for (int i = 1; i < L; ++i) {
    const u64 w = x.w[i] & spin_mask;
    a = (a ^ (a >> 15)) * 0x2C1B3C6Du + (u32)w;
    b = (b ^ (b >> 13)) * 0x297A2D39u + (u32)(w >> 32);
  }
  u32 lo = a * 0x9E3779B1u ^ b * 0x85EBCA77u;
  u32 hi = a * 0x27D4EB2Fu + b * 0x165667B1u;
  lo ^= lo >> 15;
  lo *= 0xC2B2AE3Du;
  lo ^= lo >> 13;
  hi ^= hi >> 16;
  hi *= 0x7FEB352Du;
  hi ^= hi >> 15;
  if (hi == 0xffffffffu) hi = 0xfffffffeu;
  return ((u64)hi << 32) | lo;
}
// One-word ONVs: the string fits in 32 bits once the two halves of the masked word are interleaved (every other bit of each
// half is free), and a BIJECTIVE 32-bit mix of it gives the high word; the low word is another bijection of the same value.
// Two strings with the same tag are then the same string (up to the two tag bits forced to 1), and the hash costs a
// third of the generic one.
template <>
__device__ __forceinline__ u64 hash_string<1>(const Onv<1> &x, u64 spin_mask) {
  const u64 m = x.w[0] & spin_mask;
  // the upper half moves onto the free bits of the lower half (alpha: even bits, shift left; beta: odd bits, shift right)
  const u32 f = (u32)m ^ ((spin_mask & 1ull) ? (u32)(m >> 32) << 1 : (u32)(m >> 32) >> 1);
  u32 hi = f;
  hi ^= hi >> 16;
  hi *= 0x7FEB352Du;
  hi ^= hi >> 15;
  hi *= 0x846CA68Bu;
  hi ^= hi >> 16;
  const u32 lo = hi * 0x9E3779B1u;
  if (hi == 0xffffffffu) hi = 0xfffffffeu;
  return ((u64)hi << 32) | lo;
}
template <int L>
__device__ __forceinline__ u64 hash_alpha(const Onv<L> &x) { return hash_string<L>(x, kEven); }
template <int L>
__device__ __forceinline__ u64 hash_beta(const Onv<L> &x) { return hash_string<L>(x, kOdd); }

__device__ __forceinline__ u32 hash_tag(u64 h) { return (u32)h | 3u; }
__device__ __forceinline__ bool tags_match(const uint4 &t, u32 tag) {
  return t.x == tag || t.y == tag || t.z == tag || (t.w | 1u) == tag;
}
__device__ __forceinline__ bool bucket_overflowed(const uint4 &t) { return t.w != 0u && !(t.w & 1u); }
// region descriptor packed in 64 bits: off | lg << 32; kNoRegion when the group does not exist
constexpr u64 kNoRegion = ~0ull;

__device__ __forceinline__ u64 dir_find(const DirSlot *__restrict__ dir, u32 log2_dir, u64 h) {
  const u32 mask = (1u << log2_dir) - 1u;
  u32 s = (u32)(h >> (64 - log2_dir));
  for (u32 probe = 0; probe <= mask; ++probe) {
    const uint4 e = __ldg(reinterpret_cast<const uint4 *>(dir + s));
    const u64 eh = (u64)e.x | ((u64)e.y << 32);
    if (eh == h) return (u64)e.z | ((u64)e.w << 32);
    if (eh == kDirEmpty) return kNoRegion;
    s = (s + 1) & mask;
  }
  return kNoRegion;
}

// the same probe split in two, so that a caller can put independent work between the load and its use:
// dir_first_slot / the caller's __ldg of that slot / dir_resolve (which continues the linear probing when it has to)
__device__ __forceinline__ u32 dir_first_slot(u32 log2_dir, u64 h) { return (u32)(h >> (64 - log2_dir)); }
__device__ __forceinline__ u64 dir_resolve(const DirSlot *__restrict__ dir, u32 log2_dir, u64 h, u32 s, uint4 e) {
  const u32 mask = (1u << log2_dir) - 1u;
  for (u32 probe = 0; probe <= mask; ++probe) {
    const u64 eh = (u64)e.x | ((u64)e.y << 32);
    if (eh == h) return (u64)e.z | ((u64)e.w << 32);
    if (eh == kDirEmpty) return kNoRegion;
    s = (s + 1) & mask;
    e = __ldg(reinterpret_cast<const uint4 *>(dir + s));
  }
  return kNoRegion;
}

// probe the region `desc` for the string hash h2; make_key() builds the query only when a tag matches
template <int L, typename MakeKey>
__device__ __forceinline__ long long region_probe(const u64 *__restrict__ key, const IndexView &iv, u64 desc, u64 h2,
                                                  MakeKey make_key) {
  const u32 off = (u32)desc, lg = (u32)(desc >> 32);
  const u32 mask = (1u << lg) - 1u;
  const u32 tag = hash_tag(h2);
  u32 b = lg ? (u32)(h2 >> (64 - lg)) : 0u;
  for (u32 probe = 0; probe <= mask; ++probe) {
    const uint4 t = __ldg(iv.pool + off + b);
    if (tags_match(t, tag)) {
      const Onv<L> q = make_key();
      const uint4 ids = __ldg(iv.rows + off + b);
      const u32 tg[4] = {t.x, t.y, t.z, t.w | 1u}, id[4] = {ids.x, ids.y, ids.z, ids.w};
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (tg[s] == tag) {
          const Onv<L> e = load_onv<L>(key + (long long)id[s] * L);
          if (eq_onv<L>(e, q)) return (long long)id[s];
        }
      }
    }
    if (!bucket_overflowed(t)) return -1;
    b = (b + 1) & mask;
  }
  return -1;
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

// full lookup of an arbitrary ONV through the beta-grouped directory (tables without duplicates)
template <int L>
__device__ __forceinline__ long long indexed_search(const u64 *__restrict__ key, const IndexView &iv, const Onv<L> &q) {
  const u64 desc = dir_find(iv.dir[0], iv.log2_dir, hash_beta<L>(q));
  if (desc == kNoRegion) return -1;
  return region_probe<L>(key, iv, desc, hash_alpha<L>(q), [&]() { return q; });
}

}  // namespace pynqs
