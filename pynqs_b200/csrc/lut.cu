// lut.cu -- wavefunction_lut kernels: classic binary search (K6 replacement, cuda/kernel.cu:608-680)
// and the string-grouped index (build + probe) of lut.cuh.
//
// Build, all asynchronous on the caller's stream (no host round trip, no sort):
//   1. count : every key claims / finds the directory slot of its beta string and of its alpha
//              string (64-bit CAS) and bumps the slot's key count;
//              the first key of a string also appends the slot to the list of slots in use;
//   2. carve : every slot in use takes a power-of-two run of buckets (<= 1 key per 4-slot bucket on
//              average) from the shared pool with one atomicAdd;
//   3. fill  : every key inserts (tag, row) into its two regions (32-bit CAS, linear probing
//              inside the region).
#include "eloc.cuh"
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
lut_indexed_kernel(const u64 *__restrict__ key, long long N, IndexView iv, const u64 *__restrict__ q, long long n,
                   long long *__restrict__ idx, unsigned char *__restrict__ mask) {
  const bool dup = iv.hdr->has_dup != 0;  // duplicates: reproduce the reference's probe sequence instead
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    // queries and results stream through once: evict-first, so that the index stays in L2
    Onv<L> x;
#pragma unroll
    for (int w = 0; w < L; ++w) x.w[w] = __ldcs(q + t * L + w);
    const long long r = dup ? classic_search<L>(key, N, x) : indexed_search<L>(key, iv, x);
    __stcs(idx + t, r);
    __stcs(mask + t, (unsigned char)(r >= 0));
  }
}

// Four CONSECUTIVE queries per thread (one-word ONVs).  The queries of wavefunction_lut are the rows of comb[n, M]: long runs
// of them share a beta string (alpha singles, alpha-alpha doubles, the 75-row blocks of alpha-beta doubles) or an alpha string
// (beta singles, beta-beta doubles).  A thread therefore
//   1. reads its 32 bytes of queries with two 16-byte loads (a warp reads 1 KB contiguous),
//   2. finds the region of its queries with ONE directory probe when they share the beta string, or else the alpha string
//      (then through the alpha-grouped directory; a thread that straddles the end of a run probes for every query),
//   3. issues the four tag-bucket loads together, and only then
//   4. looks at the tags (a match or an overflowed bucket goes through region_probe, which verifies against the key table),
//   5. writes the four indices with two 16-byte stores and the four mask bytes with one 4-byte store.
// The dependent chain per thread is query -> directory -> tags -> store for four queries instead of for each one, and the
// hash + directory probe of the shared string is paid once.  Same answers as lut_indexed_kernel (both directories index
// every key); tables with duplicate keys and unaligned tensors keep the one-query kernel.
__global__ void __launch_bounds__(256)
lut_batched_kernel(const u64 *__restrict__ key, long long N, IndexView iv, const u64 *__restrict__ q, long long n,
                   long long *__restrict__ idx, unsigned char *__restrict__ mask) {
  constexpr int Q = 4;
  const u32 lg_dir = iv.log2_dir;
  const long long groups = n / Q;
  const bool dup = iv.hdr->has_dup != 0;  // duplicates: reproduce the reference's probe sequence instead
  // the queries of the NEXT round are fetched before the current ones are worked on: the stream of queries comes out of
  // DRAM, and a thread that waited for its own 32 bytes each round left the memory system idle (a quarter of all stall
  // samples sat on the first use of the queries)
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  ulonglong2 n0 = make_ulonglong2(0ull, 0ull), n1 = n0;
  if (g < groups) {
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(q + Q * g);
    n0 = __ldcs(src), n1 = __ldcs(src + 1);
  }
  for (; g < groups; g += stride) {
    u64 x[Q];
    x[0] = n0.x, x[1] = n0.y, x[2] = n1.x, x[3] = n1.y;
    if (g + stride < groups) {
      const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(q + Q * (g + stride));
      n0 = __ldcs(src), n1 = __ldcs(src + 1);
    }
    long long r[Q];
    if (dup) {
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        Onv<1> o;
        o.w[0] = x[j];
        r[j] = classic_search<1>(key, N, o);
      }
    } else {
    // ---- regions ----------------------------------------------------------------------------------------------------
    // One directory probe per thread when its four queries share the beta string, or else the alpha string (then through
    // the alpha-grouped directory); threads that straddle the end of a run probe the beta directory for every query.
    u64 desc[Q];
    const u64 diff = (x[0] ^ x[1]) | (x[0] ^ x[2]) | (x[0] ^ x[3]);
    const bool same_b = (diff & kOdd) == 0ull, same_a = (diff & kEven) == 0ull;
    const bool by_alpha = !same_b && same_a;  // the whole thread goes through the alpha-grouped directory
    const bool mixed = !same_b && !same_a;
    // the end of a run inside the thread's four queries (nearly every warp has such a thread: the runs are 75 rows long):
    // ONE more probe, for the beta string of the last query, covers the queries of the second run; what matches neither
    // (three runs in four queries: the seams between excitation classes) is probed one by one further down
    const DirSlot *dir = by_alpha ? iv.dir[1] : iv.dir[0];
    Onv<1> o;
    o.w[0] = x[0];
    const u64 h0 = hash_string<1>(o, by_alpha ? kEven : kOdd);
    const u32 s0 = dir_first_slot(lg_dir, h0);
    const uint4 e0 = __ldg(reinterpret_cast<const uint4 *>(dir + s0));  // (its use comes after the hashes below)
    u64 h3 = 0ull;
    u32 s3 = 0u;
    uint4 e3 = make_uint4(0u, 0u, 0u, 0u);
    if (mixed) {
      o.w[0] = x[Q - 1];
      h3 = hash_beta<1>(o);
      s3 = dir_first_slot(lg_dir, h3);
      e3 = __ldg(reinterpret_cast<const uint4 *>(iv.dir[0] + s3));
    }
    // ---- hashes of the OTHER string of every query (independent of the directory: they fill its latency) -----------------
    const u64 other = by_alpha ? kOdd : kEven;  // the string the region is hashed by
    u32 tag[Q], hiw[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      o.w[0] = x[j];
      const u64 h2 = hash_string<1>(o, other);
      tag[j] = hash_tag(h2);
      hiw[j] = (u32)(h2 >> 32);
    }
    {
      const u64 d0 = dir_resolve(dir, lg_dir, h0, s0, e0);
#pragma unroll
      for (int j = 0; j < Q; ++j) desc[j] = d0;
      if (mixed) {
        const u64 d3 = dir_resolve(iv.dir[0], lg_dir, h3, s3, e3);
        desc[Q - 1] = d3;
#pragma unroll
        for (int j = 1; j < Q - 1; ++j) {
          if (((x[j] ^ x[Q - 1]) & kOdd) == 0ull) {
            desc[j] = d3;
          } else if (((x[j] ^ x[0]) & kOdd) != 0ull) {
            o.w[0] = x[j];
            desc[j] = dir_find(iv.dir[0], lg_dir, hash_beta<1>(o));
          }
        }
      }
    }
    // ---- tag buckets: four loads in flight ----------------------------------------------------------------------------
    uint4 t[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      t[j] = make_uint4(0u, 0u, 0u, 0u);
      const u32 off = (u32)desc[j], lg = (u32)(desc[j] >> 32);
      if ((int)lg >= 0) {  // (kNoRegion has all bits set; a real region has lg <= 31)
        // top lg bits of the high hash word = h2 >> (64 - lg), and 0 for a one-bucket region, in one funnel shift
        const u32 bkt = __funnelshift_l(hiw[j], 0u, lg);
        t[j] = __ldg(iv.pool + off + bkt);
      }
    }
    // ---- results ---------------------------------------------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      r[j] = -1;
      if (tags_match(t[j], tag[j]) || bucket_overflowed(t[j])) {  // rare: a hit, a tag collision or a full bucket
        o.w[0] = x[j];
        const u64 h2 = hash_string<1>(o, other);
        r[j] = region_probe<1>(key, iv, desc[j], h2, [&]() { return o; });
      }
    }
    }
    longlong2 *dst = reinterpret_cast<longlong2 *>(idx + Q * g);
    __stcs(dst, make_longlong2(r[0], r[1]));
    __stcs(dst + 1, make_longlong2(r[2], r[3]));
    const u32 m4 = (u32)(r[0] >= 0) | ((u32)(r[1] >= 0) << 8) | ((u32)(r[2] >= 0) << 16) | ((u32)(r[3] >= 0) << 24);
    __stcs(reinterpret_cast<unsigned int *>(mask + Q * g), m4);
  }
}

// header + both directories: every slot {h = empty, off = 0, count = 0}
__global__ void __launch_bounds__(256) index_init_kernel(HashHeader *hdr, uint4 *dir, u32 log2_dir, u64 N, u32 pool_buckets) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    hdr->log2_dir = log2_dir;
    hdr->has_dup = 0;
    hdr->n_keys = N;
    hdr->cursor = 0;
    hdr->pool_buckets = pool_buckets;
    hdr->n_claimed = 0;
  }
  const size_t slots = (size_t)2 << log2_dir;
  const uint4 empty = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += (size_t)gridDim.x * blockDim.x) dir[i] = empty;
}

__device__ __forceinline__ u32 dir_claim(DirSlot *dir, u32 log2_dir, u64 h, bool &first) {
  const u32 mask = (1u << log2_dir) - 1u;
  u32 s = (u32)(h >> (64 - log2_dir));
  for (;;) {
    const u64 prev = atomicCAS(reinterpret_cast<unsigned long long *>(&dir[s].h), (unsigned long long)kDirEmpty, (unsigned long long)h);
    first = prev == kDirEmpty;
    if (first || prev == h) return s;
    s = (s + 1) & mask;
  }
}

// append `entry` to the list for the lanes with `take` set: one atomicAdd per warp
__device__ __forceinline__ void list_append(u32 *counter, u32 *list, bool take, u32 entry) {
  const unsigned m = __ballot_sync(0xffffffffu, take);
  if (m == 0) return;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  u32 base = 0;
  if (lane == leader) base = atomicAdd(counter, (u32)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (take) list[base + __popc(m & ((1u << lane) - 1u))] = entry;
}

template <int L>
__global__ void __launch_bounds__(256)
index_count_kernel(const u64 *__restrict__ key, long long N, HashHeader *hdr, DirSlot *dirB, DirSlot *dirA, u32 *slotB, u32 *slotA,
                   u32 *claimed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < N;  // no early return: the list appends are warp-wide
  bool firstB = false, firstA = false;
  u32 sb = 0, sa = 0;
  if (valid) {
    const Onv<L> x = load_onv<L>(key + i * L);
    if (i > 0) {
      const Onv<L> prev = load_onv<L>(key + (i - 1) * L);
      if (eq_onv<L>(prev, x)) atomicExch(&hdr->has_dup, 1u);
    }
    const u32 lg = hdr->log2_dir;
    sb = dir_claim(dirB, lg, hash_beta<L>(x), firstB);
    atomicAdd(&dirB[sb].lg, 1u);
    slotB[i] = sb;
    sa = dir_claim(dirA, lg, hash_alpha<L>(x), firstA);
    atomicAdd(&dirA[sa].lg, 1u);
    slotA[i] = sa;
  }
  list_append(&hdr->n_claimed, claimed, firstB, sb);
  list_append(&hdr->n_claimed, claimed, firstA, sa | 0x80000000u);  // bit 31: the alpha-grouped directory
}

// count -> region: buckets = pow2ceil(count) (< 2 count, so the pool of 4N buckets suffices): at most
// one key per bucket on average, which keeps overflowed buckets (second probes) below 1 %
__global__ void __launch_bounds__(256) index_carve_kernel(HashHeader *hdr, DirSlot *dirB, DirSlot *dirA, const u32 *__restrict__ claimed) {
  const u32 n = hdr->n_claimed;
  for (u32 t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const u32 e = claimed[t];
    DirSlot *d = (e >> 31) ? dirA + (e & 0x7fffffffu) : dirB + e;
    const u32 want = d->lg;
    u32 lg = 0;
    while ((1u << lg) < want) ++lg;
    d->off = atomicAdd(&hdr->cursor, 1u << lg);
    d->lg = lg;
  }
}

__device__ __forceinline__ void region_insert(u32 *tags, u32 *rows, const DirSlot &d, u64 h2, u32 row) {
  const u32 mask = (1u << d.lg) - 1u;
  const u32 tag = hash_tag(h2);
  u32 b = d.lg ? (u32)(h2 >> (64 - d.lg)) : 0u;
  for (u32 probe = 0; probe <= mask; ++probe) {
    const size_t base = 4 * (size_t)(d.off + b);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (atomicCAS(&tags[base + s], 0u, tag) == 0u) {
        rows[base + s] = row;
        return;
      }
    }
    atomicAnd(&tags[base + 3], ~1u);  // full: raise the overflow flag and move on
    b = (b + 1) & mask;
  }
}

template <int L>
__global__ void __launch_bounds__(256)
index_fill_kernel(const u64 *__restrict__ key, long long N, const DirSlot *__restrict__ dirB, const DirSlot *__restrict__ dirA,
                  u32 *tags, u32 *rows, const u32 *__restrict__ slotB, const u32 *__restrict__ slotA) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const Onv<L> x = load_onv<L>(key + i * L);
  const u64 ha = hash_alpha<L>(x), hb = hash_beta<L>(x);
  region_insert(tags, rows, dirB[slotB[i]], ha, (u32)i);  // grouped by beta string, hashed by alpha string
  region_insert(tags, rows, dirA[slotA[i]], hb, (u32)i);  // grouped by alpha string, hashed by beta string
}

static inline unsigned grid_for(long long n, int threads, long long cap) {
  long long want = (n + threads - 1) / threads;
  if (want < 1) want = 1;
  return (unsigned)(want < cap ? want : cap);
}

long long hash_workspace_bytes(long long N) { return index_layout(N).total; }

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
  if (N >= (1LL << 31)) {
    set_error("lookup index supports fewer than 2^31 keys (got %lld)", N);
    return 1;
  }
  const IndexLayout l = index_layout(N);
  if (ws_bytes < l.total) {
    set_error("index workspace too small: %lld < %lld bytes", ws_bytes, l.total);
    return 4;
  }
  char *b = static_cast<char *>(ws);
  HashHeader *hdr = reinterpret_cast<HashHeader *>(b);
  DirSlot *dirB = reinterpret_cast<DirSlot *>(b + l.dir_off[0]);
  DirSlot *dirA = reinterpret_cast<DirSlot *>(b + l.dir_off[1]);
  u32 *pool = reinterpret_cast<u32 *>(b + l.pool_off);
  u32 *rows = reinterpret_cast<u32 *>(b + l.idx_off);
  u32 *slotB = reinterpret_cast<u32 *>(b + l.scratch_off);
  u32 *slotA = slotB + N;
  u32 *claimed = slotA + N;
  if (cudaMemsetAsync(pool, 0, (size_t)(l.idx_off - l.pool_off), st) != cudaSuccess) return check_launch("index memset");
  index_init_kernel<<<148 * 8, 256, 0, st>>>(hdr, reinterpret_cast<uint4 *>(dirB), l.log2_dir, (u64)N, (u32)l.pool_buckets);
  count_launch();
  if (N > 0) {
    const unsigned blocks = (unsigned)((N + 255) / 256);
    switch (L) {
      case 1: index_count_kernel<1><<<blocks, 256, 0, st>>>(key, N, hdr, dirB, dirA, slotB, slotA, claimed); break;
      case 2: index_count_kernel<2><<<blocks, 256, 0, st>>>(key, N, hdr, dirB, dirA, slotB, slotA, claimed); break;
      case 3: index_count_kernel<3><<<blocks, 256, 0, st>>>(key, N, hdr, dirB, dirA, slotB, slotA, claimed); break;
      default: set_error("unsupported ONV length L=%d", L); return 1;
    }
    count_launch();
    index_carve_kernel<<<grid_for(2 * N, 256, 148LL * 4), 256, 0, st>>>(hdr, dirB, dirA, claimed);
    count_launch();
    switch (L) {
      case 1: index_fill_kernel<1><<<blocks, 256, 0, st>>>(key, N, dirB, dirA, pool, rows, slotB, slotA); break;
      case 2: index_fill_kernel<2><<<blocks, 256, 0, st>>>(key, N, dirB, dirA, pool, rows, slotB, slotA); break;
      default: index_fill_kernel<3><<<blocks, 256, 0, st>>>(key, N, dirB, dirA, pool, rows, slotB, slotA); break;
    }
    count_launch();
  }
  return check_launch("index build");
}

int launch_lut_hashed(const u64 *key, long long N, const u64 *q, long long n, int L, const void *ws, long long *idx,
                      unsigned char *mask, cudaStream_t st) {
  if (n == 0) return 0;
  const IndexView iv = index_view(ws, N);
  // one-word ONVs, 16-byte aligned tensors: four consecutive queries per thread; the (up to three) last queries and every
  // other case go through the one-query kernel.  Both kernels fall back to the reference's probe sequence by themselves when
  // the build found duplicate keys.
  long long done = 0;
  const bool aligned = (reinterpret_cast<uintptr_t>(q) & 15u) == 0 && (reinterpret_cast<uintptr_t>(idx) & 15u) == 0 &&
                       (reinterpret_cast<uintptr_t>(mask) & 3u) == 0;
  if (L == 1 && aligned && n >= 4 && eloc_tuning().lut_pipeline) {
    const long long groups = n / 4;
    lut_batched_kernel<<<grid_for(groups, 256, 148LL * 64), 256, 0, st>>>(key, N, iv, q, groups * 4, idx, mask);
    count_launch();
    if (int rc = check_launch("lut_batched_kernel")) return rc;
    done = groups * 4;
    if (done == n) return 0;
  }
  const long long rest = n - done;
  const unsigned blocks = grid_for(rest, 256, 148LL * 64);
  switch (L) {
    case 1: lut_indexed_kernel<1><<<blocks, 256, 0, st>>>(key, N, iv, q + done, rest, idx + done, mask + done); break;
    case 2: lut_indexed_kernel<2><<<blocks, 256, 0, st>>>(key, N, iv, q, n, idx, mask); break;
    case 3: lut_indexed_kernel<3><<<blocks, 256, 0, st>>>(key, N, iv, q, n, idx, mask); break;
    default: set_error("unsupported ONV length L=%d", L); return 1;
  }
  count_launch();
  return check_launch("lut_indexed_kernel");
}

}  // namespace pynqs
