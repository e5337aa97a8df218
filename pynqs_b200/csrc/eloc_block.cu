// eloc_block.cu -- the scan of the one-pass local energy for samples that SHARE a beta string.
//
// The per-sample kernel (eloc_scan.cu) walks, for every sample, the ~noB*nvB + 2 groups of table keys that can hold a
// connected determinant.  All but one of those groups depend on the sample's BETA string only, and in a VMC sample
// set many samples carry the same beta string (10^6 Fe2S2 samples: at most C(20,15) = 15 504 beta strings, so >= 64
// samples per string).  Here the samples are first grouped by beta string (three small kernels, no sort: the order
// of the samples inside a group is irrelevant), and one warp then takes a TILE of up to 32 samples with the same
// beta string -- one sample per lane -- and walks the shared groups once for all of them:
//
//   * the folded alpha strings of the groups' keys are copied into shared memory with cp.async (coalesced, all
//     copies of a fill in flight at once) and read back as broadcast LDS.128;
//   * per key and per 32 samples the test is three instructions (XOR, POPC, compare-and-OR into a predicate);
//     a lane only records WHICH block of 8 keys held a hit, the block is re-examined when the queue is flushed;
//   * bucket bounds, hashing, duplicate-bucket detection and the copy are paid once per tile, not per sample.
//
// Compared with the per-sample kernel (16 instructions per 32 key tests plus ~2 000 per sample of set-up) this is
// ~4 instructions per 32 tests and ~100 per sample of set-up.  Groups with fewer than `block_min_group` samples stay
// with the per-sample kernel (a lane-per-sample walk with one lane is slower than a lane-per-key walk).
//
// The sample's own-alpha group (beta singles, beta-beta doubles) depends on the sample, not on the tile: it is walked
// per sample with one key per lane, as in the per-sample kernel.
//
// Output: the same hit lists (HitRun records + hit words) the per-sample kernel writes; eloc_eval_kernel reads both.
// One-word ONVs only (folded 32-bit strings); tables with duplicate keys are handed to the evaluation kernel's
// reference route like everywhere else.
#include "eloc.cuh"

namespace pynqs {

constexpr int kBlkWarps = 4;     // warps per CTA; every warp works on its own tile
#ifndef PYNQS_STAGE
#define PYNQS_STAGE 512
#endif
#ifndef PYNQS_QCAP
#define PYNQS_QCAP 32
#endif
constexpr int kStage = PYNQS_STAGE;  // folded strings staged per warp and fill
constexpr int kQCap = PYNQS_QCAP;    // queue entries per lane (16-bit: index of a block of 8 staged keys)
constexpr int kMaxGroups = 260;  // beta singles of a one-word ONV (<= 16 * 16) + the own group
constexpr int kFirstBlock = 64;  // hit-buffer block sizes of a sample: 64, 128, ... (one HitRun each)

struct BlockWarpSmem {
  u32 stage[kStage];
  unsigned short queue[kQCap * 32];  // [entry][lane]
  uint2 rng[kMaxGroups];             // bucket [first, end) of every group; group sB = the tile's own beta string
  u32 goff[kMaxGroups];              // first block (of 8 keys) of every group among the blocks of its walk
  u32 blkpos[kStage / 8];            // position in the grouped copy of the first key of every staged block of 8
  unsigned char blkcnt[kStage / 8];  // keys in the block (the rest is padding)
  unsigned char bpos[32];
};
constexpr int kDupSet = 512;  // open-addressing set of the bucket starts of a walk (lives in the idle queue)
static_assert(kDupSet * 4 <= kQCap * 32 * 2 && kDupSet <= kStage && kDupSet >= 2 * 256, "duplicate-bucket set: twice the largest number of groups");

__host__ __device__ inline u32 sample_log2_buckets(long long n) {
  u32 lg = 10;
  while ((1LL << lg) < 4 * n && lg < 24) ++lg;
  return lg;
}

// ---- grouping pass ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sample_count_kernel(const u64 *__restrict__ bra, long long n, u32 shift, u32 *__restrict__ bcnt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Onv<1> x;
  x.w[0] = bra[i];
  atomicAdd(bcnt + (string_hash32<1>(x, kOdd) >> shift), 1u);
}

// one thread per bucket: a range of the slot array for its samples and, when there are enough of them, tiles of
// 17..32 samples (balanced); small buckets go to the back of the slot array, for the per-sample kernel
__global__ void __launch_bounds__(256)
sample_alloc_kernel(u32 *__restrict__ bcnt, u32 *__restrict__ bbase, u32 nbuckets, ElocCounters *ctr, uint2 *__restrict__ tiles, u32 min_group,
                    u32 n) {
  const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbuckets) return;
  const u32 c = bcnt[b];
  if (c == 0) return;
  bcnt[b] = 0;  // becomes the fill counter of the scatter pass
  if (c < min_group) {
    bbase[b] = n - c - atomicAdd(&ctr->n_single, c);
    return;
  }
  const u32 base = atomicAdd(&ctr->slot_front, c);
  bbase[b] = base;
  const u32 nt = (c + 31u) >> 5, lo = c / nt, extra = c - lo * nt;  // `extra` tiles of lo + 1 samples, the rest lo
  u32 t = atomicAdd(&ctr->tile_count, nt), at = base;
  for (u32 k = 0; k < nt; ++k) {
    const u32 sz = lo + (k < extra ? 1u : 0u);
    tiles[t + k] = make_uint2(at, sz);
    at += sz;
  }
}

__global__ void __launch_bounds__(256)
sample_scatter_kernel(const u64 *__restrict__ bra, long long n, u32 shift, u32 *__restrict__ bfill, const u32 *__restrict__ bbase,
                      u32 *__restrict__ slots) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Onv<1> x;
  x.w[0] = bra[i];
  const u32 b = string_hash32<1>(x, kOdd) >> shift;
  slots[bbase[b] + atomicAdd(bfill + b, 1u)] = (u32)i;
}

// ---- the block kernel -------------------------------------------------------------------------------------------------
struct BlockGeom {
  int sorb, noA, noB, nvB;
  u32 never;  // stands in for the alpha string of an idle lane: at distance > 4 from every key and pad
  u32 pad;    // pads the staged blocks: at distance > 4 from every alpha string
  int pow2;   // noA > 4: the alpha-beta test is "a & ~k has at most one bit" (no POPC), which pad = 0 never passes
};

struct LaneHits {  // where a lane's sample keeps its hits in the global buffer
  u32 base, fill, cap, nrun;
  u32 end;  // run records [.., end) of the sample belong to this part of the tile
  bool over;
};

// what the rarely taken paths of a tile need (kept in local memory; the inner loop never touches it)
struct TileCtx {
  BlockWarpSmem *S;
  const u64 *keysB;
  u32 *self_pos, *hits;
  HitRun *my_runs;
  ElocCounters *ctr;
  u64 x;
  u32 a, sid, hit_cap;
  int run_stride;
};

// a lane's current block is full (or there is none yet): close it and take the next, twice as large.  The bookkeeping
// travels by value so that it stays in registers on the caller's side (every emitted hit reads it).
__device__ __noinline__ LaneHits next_hit_block(const TileCtx &c, LaneHits h) {
  if (h.cap) {
    c.my_runs[h.nrun] = HitRun{h.base, h.fill};
    ++h.nrun;
  }
  const u32 ncap = h.cap ? 2u * h.cap : (u32)kFirstBlock;
  if (h.nrun == h.end) {
    h.over = true;
    return h;
  }
  const u32 off = atomicAdd(&c.ctr->hit_cursor, ncap);
  if (off > c.hit_cap || ncap > c.hit_cap - off) {
    h.over = true;
    return h;
  }
  h.base = off;
  h.cap = ncap;
  h.fill = 0;
  return h;
}

__device__ __forceinline__ void emit_hit(const TileCtx &c, LaneHits &h, u32 v) {
  if (h.over) return;
  if (h.fill == h.cap) {
    h = next_hit_block(c, h);
    if (h.over) return;
  }
  c.hits[h.base + h.fill++] = v;
}

__device__ __forceinline__ void cp_async4(u32 *smem_dst, const u32 *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// the queue holds blocks (8 staged keys) that contained a hit: look at each key again and emit the hits.  The lanes
// stay together: entry i of every lane is examined at the same time, then the hits are emitted bit by bit
template <bool OWN>
__device__ __noinline__ LaneHits flush_blocks(const TileCtx &c, u32 qn, LaneHits h) {
  BlockWarpSmem &S = *c.S;
  const int lane = threadIdx.x & 31;
  const u32 a = c.a;
  const u32 mx = __reduce_max_sync(0xffffffffu, qn);
  for (u32 i = 0; i < mx; ++i) {
    u32 m = 0, gpos = 0;
    if (i < qn) {
      const u32 j = S.queue[i * 32 + lane];
      gpos = S.blkpos[j >> 3];
      const uint4 k0 = *reinterpret_cast<const uint4 *>(&S.stage[j]), k1 = *reinterpret_cast<const uint4 *>(&S.stage[j + 4]);
      const u32 k[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const u32 d = (u32)__popc(k[u] ^ a);
        if (OWN) {
          if (d == 0u) {
            if (c.keysB[gpos + u] == c.x) c.self_pos[c.sid] = gpos + u;  // the sample itself (checked on the full key)
          } else if (d == 2u || d == 4u) {
            m |= 1u << u;
          }
        } else if (d == 2u) {
          m |= 1u << u;
        }
      }
    }
    while (__any_sync(0xffffffffu, m != 0u)) {
      if (m) {
        const int u = __ffs(m) - 1;
        m &= m - 1u;
        emit_hit(c, h, (gpos + (u32)u) | (OWN ? kHitOwn : 0u));
      }
    }
  }
  return h;
}

// test the staged keys [0, cnt) (cnt a multiple of 16) against the lane's alpha string: 3 instructions per key
// (XOR, POPC, compare-and-accumulate into a predicate), one queue entry per block of 8 keys with a hit
// (POW2: a and k have the same number of bits, so they differ in one orbital pair iff a & ~k is a single bit;
//  v & (v - 1) == 0 also lets k == a through, which the exact test of flush_blocks drops.  ALU / FMA pipes instead
//  of the quarter-rate POPC pipe: 7.0 instead of 8.8 cycles per key and scheduler, profiles/micro/test_rate.cu)
__device__ __forceinline__ bool one_pair(u32 a, u32 k) {
  const u32 v = a & ~k;
  return (v & (v - 1u)) == 0u;
}

template <bool OWN, bool POW2>
__device__ __forceinline__ void test_stage(const TileCtx &c, LaneHits &h, BlockWarpSmem &S, u32 cnt, u32 a) {
  const int lane = threadIdx.x & 31;
  const uint4 *st4 = reinterpret_cast<const uint4 *>(S.stage);
  char *qbase = reinterpret_cast<char *>(&S.queue[lane]);
  u32 qoff = 0;  // 64 * number of queued entries
  for (u32 j = 0; j < cnt; j += 16) {
    const uint4 k0 = st4[(j >> 2)], k1 = st4[(j >> 2) + 1], k2 = st4[(j >> 2) + 2], k3 = st4[(j >> 2) + 3];
    bool h0, h1;
    if (OWN) {
      h0 = (__popc(k0.x ^ a) <= 4) | (__popc(k0.y ^ a) <= 4) | (__popc(k0.z ^ a) <= 4) | (__popc(k0.w ^ a) <= 4) | (__popc(k1.x ^ a) <= 4) |
           (__popc(k1.y ^ a) <= 4) | (__popc(k1.z ^ a) <= 4) | (__popc(k1.w ^ a) <= 4);
      h1 = (__popc(k2.x ^ a) <= 4) | (__popc(k2.y ^ a) <= 4) | (__popc(k2.z ^ a) <= 4) | (__popc(k2.w ^ a) <= 4) | (__popc(k3.x ^ a) <= 4) |
           (__popc(k3.y ^ a) <= 4) | (__popc(k3.z ^ a) <= 4) | (__popc(k3.w ^ a) <= 4);
    } else if (POW2) {
      h0 = one_pair(a, k0.x) | one_pair(a, k0.y) | one_pair(a, k0.z) | one_pair(a, k0.w) | one_pair(a, k1.x) | one_pair(a, k1.y) |
           one_pair(a, k1.z) | one_pair(a, k1.w);
      h1 = one_pair(a, k2.x) | one_pair(a, k2.y) | one_pair(a, k2.z) | one_pair(a, k2.w) | one_pair(a, k3.x) | one_pair(a, k3.y) |
           one_pair(a, k3.z) | one_pair(a, k3.w);
    } else {
      h0 = (__popc(k0.x ^ a) == 2) | (__popc(k0.y ^ a) == 2) | (__popc(k0.z ^ a) == 2) | (__popc(k0.w ^ a) == 2) | (__popc(k1.x ^ a) == 2) |
           (__popc(k1.y ^ a) == 2) | (__popc(k1.z ^ a) == 2) | (__popc(k1.w ^ a) == 2);
      h1 = (__popc(k2.x ^ a) == 2) | (__popc(k2.y ^ a) == 2) | (__popc(k2.z ^ a) == 2) | (__popc(k2.w ^ a) == 2) | (__popc(k3.x ^ a) == 2) |
           (__popc(k3.y ^ a) == 2) | (__popc(k3.z ^ a) == 2) | (__popc(k3.w ^ a) == 2);
    }
    if (h0) {
      *reinterpret_cast<unsigned short *>(qbase + qoff) = (unsigned short)j;
      qoff += 64u;
    }
    if (h1) {
      *reinterpret_cast<unsigned short *>(qbase + qoff) = (unsigned short)(j + 8u);
      qoff += 64u;
    }
    if (__any_sync(0xffffffffu, qoff > 64u * (u32)(kQCap - 2))) {
      h = flush_blocks<OWN>(c, qoff >> 6, h);
      qoff = 0;
    }
  }
  h = flush_blocks<OWN>(c, qoff >> 6, h);
}

// Walk the groups [q_lo, q_hi) of the current beta string: S.goff[q_lo .. q_hi] holds the exclusive prefix of their
// sizes in blocks of 8 keys.  Fill after fill: block descriptors (lanes over groups), copy (8 lanes per block, cp.async),
// test.
template <bool OWN>
__device__ __forceinline__ void walk_groups(TileCtx &c, LaneHits &h, BlockWarpSmem &S, const BlockGeom &g, const u32 *__restrict__ halfB,
                                            int q_lo, int q_hi, u32 a) {
  const int lane = threadIdx.x & 31;
  const u32 total = S.goff[q_hi];
  for (u32 fb = 0; fb < total; fb += (u32)(kStage / 8)) {
    const u32 nblk = total - fb < (u32)(kStage / 8) ? total - fb : (u32)(kStage / 8);
    for (int q = q_lo + lane; q < q_hi; q += 32) {
      const u32 g0 = S.goff[q], g1 = S.goff[q + 1];
      const u32 lo = g0 > fb ? g0 : fb, hi = g1 < fb + nblk ? g1 : fb + nblk;
      if (lo < hi) {
        const uint2 r = S.rng[q];
        for (u32 b = lo; b < hi; ++b) {
          const u32 first = r.x + 8u * (b - g0), left = r.y - first;
          S.blkpos[b - fb] = first;
          S.blkcnt[b - fb] = (unsigned char)(left < 8u ? left : 8u);
        }
      }
    }
    __syncwarp();
    const u32 nblk2 = (nblk + 1u) & ~1u;  // the test loop takes 16 keys at a time
    const u32 sub = (u32)lane & 7u;
    // 8 lanes per block, 4 blocks per pass, 4 passes in flight (the descriptors come out of shared memory: without the
    // unrolling every pass would wait for its own two loads)
    for (u32 b0 = (u32)lane >> 3; b0 < nblk2; b0 += 16) {
      u32 cnt[4], pos[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const u32 b = b0 + 4u * (u32)u;
        const bool in = b < nblk;
        cnt[u] = in ? (u32)S.blkcnt[b] : 0u;
        pos[u] = in ? S.blkpos[b] : 0u;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const u32 b = b0 + 4u * (u32)u;
        if (b < nblk2) {
          u32 *dst = &S.stage[8u * b + sub];
          if (sub < cnt[u]) cp_async4(dst, halfB + pos[u] + sub);
          else *dst = g.pad;
        }
      }
    }
    cp_async_wait_all();
    __syncwarp();
    c.a = a;
    if (OWN) test_stage<true, false>(c, h, S, 8u * nblk2, a);
    else if (g.pow2) test_stage<false, true>(c, h, S, 8u * nblk2, a);
    else test_stage<false, false>(c, h, S, 8u * nblk2, a);
    __syncwarp();
  }
}

// SPLIT: the alpha-beta groups of a tile are split over several warps (see `parts` below); a template parameter so that
// the whole-tile kernel of large calls carries none of the bookkeeping.
template <bool SPLIT>
__global__ void __launch_bounds__(kBlkWarps * 32)
eloc_block_kernel(const u64 *__restrict__ bra, const u32 *__restrict__ slots, const uint2 *__restrict__ tiles, ElocCounters *ctr, GroupView gv,
                  HitRun *__restrict__ runs, u32 *__restrict__ run_cnt, int run_stride, u32 *__restrict__ hits, u32 *__restrict__ self_pos,
                  u32 hit_cap, BlockGeom g, u32 force_parts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  BlockWarpSmem &S = reinterpret_cast<BlockWarpSmem *>(smem_raw)[warp];
  const u32 table_has_dup = __ldg(&gv.hdr->has_dup);
  const u32 ntiles = ctr->tile_count;  // final: written by the grouping pass, which finished before this kernel started
  const int sB = g.noB * g.nvB;
  // Few tiles (a rank's share of the samples at 8 GPUs; the launcher decides by the number of samples): with ~1.2 tiles per
  // resident warp the launch takes the time of two tiles.  The alpha-beta groups of a tile are then split into `parts` ranges that different warps walk (part 0 also takes
  // the tile's own group and the samples' own-alpha groups); every part writes its hits to run records of its own
  // (run_stride / parts each), so nothing is shared between the parts but the read-only table.  Every part repeats the
  // tile's set-up (~12 % of a tile), so two parts is where it stops paying: measured for 125 000 samples 0.700 / 0.656 /
  // 0.738 ms with 1 / 2 / 4 parts (profiles/micro/eloc_slice.py).
  u32 parts = 1;
  if (SPLIT) {
    parts = force_parts ? force_parts : 2u;
    if ((u32)run_stride < 2u * parts) parts = 1;
  }
  const u32 runs_per_part = (u32)run_stride / parts;
  const u32 nitems = ntiles * parts;
  const u32 *__restrict__ halfB = gv.half[0];
  const u32 *__restrict__ halfA = gv.half[1];
  const u32 all_beta = fold_beta(g.sorb >= 64 ? ~0ull : ((1ull << g.sorb) - 1ull));
  for (;;) {
    u32 t = 0;
    if (lane == 0) t = atomicAdd(&ctr->tile_next, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= nitems) break;
    const u32 part = t / ntiles;  // part-major: the (heavier) parts 0 are handed out first
    t -= part * ntiles;
    const uint2 tile = tiles[t];
    const bool active = (u32)lane < tile.y;
    const u32 sid = active ? slots[tile.x + lane] : 0u;
    const u64 x = active ? bra[sid] : 0ull;
    const u32 fa = fold_alpha(x), fb = fold_beta(x);
    TileCtx c;
    c.S = &S;
    c.keysB = gv.keys[0];
    c.self_pos = self_pos;
    c.hits = hits;
    c.my_runs = runs + (size_t)sid * run_stride;
    c.ctr = ctr;
    c.x = x;
    c.sid = sid;
    c.hit_cap = hit_cap;
    c.run_stride = run_stride;
    LaneHits h = {0u, 0u, 0u, part * runs_per_part, (part + 1u) * runs_per_part, table_has_dup != 0u};
    const int q_part_lo = (int)((u32)sB * part / parts), q_part_hi = (int)((u32)sB * (part + 1u) / parts);
    u32 remaining = table_has_dup ? 0u : __ballot_sync(0xffffffffu, active);
    // a tile holds the samples of one bucket of the grouping pass: normally one beta string, after a hash collision several
    while (remaining) {
      const int leader = __ffs(remaining) - 1;
      const u32 fbL = __shfl_sync(0xffffffffu, fb, leader);
      const bool mine = active && fb == fbL;
      remaining &= ~__ballot_sync(0xffffffffu, mine);
      const u32 a = mine ? fa : g.never;
      // ---- the groups of this beta string: positions of its occupied / virtual orbitals, then one bucket per single ------
      __syncwarp();
      {
        const u32 bit = 1u << lane, below = bit - 1u;
        if (fbL & bit) S.bpos[__popc(fbL & below)] = (unsigned char)lane;
        else if (all_beta & bit) S.bpos[g.noB + __popc(all_beta & ~fbL & below)] = (unsigned char)lane;
      }
      __syncwarp();
      for (int q = lane; q <= sB; q += 32) {
        u32 f = fbL;
        if (q < sB) {
          const int pb = q / g.noB, hb = q - pb * g.noB;
          f ^= (1u << S.bpos[hb]) ^ (1u << S.bpos[g.noB + pb]);
        }
        Onv<1> y;
        y.w[0] = unfold_beta_word(f);
        const u32 *stp = gv.start[0] + group_bucket<1>(y, 0, gv.shift);
        S.rng[q] = make_uint2(__ldg(stp), __ldg(stp + 1));
      }
      __syncwarp();
      // the own group first (other acceptance test): the keys that carry the tile's beta string itself
      if (lane == 0) {
        const uint2 r = S.rng[sB];
        S.goff[sB] = 0u;
        S.goff[sB + 1] = (r.y - r.x + 7u) >> 3;
      }
      __syncwarp();
      if (part == 0u) walk_groups<true>(c, h, S, g, halfB, sB, sB + 1, a);
      // two singles whose strings share a bucket: the folded test cannot tell them apart, so the bucket is walked for one
      // of them only -- the one with the SMALLEST group number, so that every part of a split tile drops the same ones.
      // Buckets are told apart by their first position; the set lives in the (now idle) hit queue, the smallest group number
      // of every bucket in the (now idle) stage buffer, a group's slot of the set in goff (recomputed below)
      u32 *dupset = reinterpret_cast<u32 *>(S.queue);
      u32 *minq = S.stage;
      for (int i = lane; i < kDupSet; i += 32) {
        dupset[i] = 0xffffffffu;
        minq[i] = 0xffffffffu;
      }
      __syncwarp();
      for (int q = lane; q < sB; q += 32) {
        const uint2 r = S.rng[q];
        if (r.x == r.y) continue;
        u32 slot = (r.x * 0x9E3779B1u) >> (32 - 9);
        for (;;) {
          const u32 old = atomicCAS(&dupset[slot], 0xffffffffu, r.x);
          if (old == 0xffffffffu || old == r.x) break;
          slot = (slot + 1u) & (u32)(kDupSet - 1);
        }
        atomicMin(&minq[slot], (u32)q);
        S.goff[q] = slot;
      }
      __syncwarp();
      for (int q = lane; q < sB; q += 32) {
        const uint2 r = S.rng[q];
        if (r.x != r.y && minq[S.goff[q]] != (u32)q) S.rng[q] = make_uint2(0u, 0u);
      }
      __syncwarp();
      // first block of every group: exclusive prefix over the groups' sizes in blocks of 8
      {
        u32 running = 0;
        for (int q0 = 0; q0 < sB; q0 += 32) {
          const int q = q0 + lane;
          u32 nb = 0;
          if (q >= q_part_lo && q < q_part_hi) {  // (the buckets of the other parts' groups count as empty here)
            const uint2 r = S.rng[q];
            nb = (r.y - r.x + 7u) >> 3;
          }
          u32 incl = nb;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const u32 up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
          }
          if (q < sB) S.goff[q] = running + incl - nb;
          running += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) S.goff[sB] = running;
      }
      __syncwarp();
      walk_groups<false>(c, h, S, g, halfB, 0, sB, a);
    }
    // ---- every sample's own-alpha group (beta singles, beta-beta doubles): the group depends on the sample, so every lane
    // walks ITS bucket of the alpha-grouped copy, four independent loads in flight (a lane reads consecutive words: its lines
    // stay in L1; the 32 lanes touch 32 lines per load, which at ~3 loads per sample and key-chunk is nowhere near a limit).
    // No shuffles, no serial dependence between samples.
    if (!table_has_dup && part == 0u) {
      uint2 rA = make_uint2(0u, 0u);
      if (active) {
        Onv<1> y;
        y.w[0] = x;
        const u32 *stp = gv.start[1] + group_bucket<1>(y, 1, gv.shift);
        rA = make_uint2(__ldg(stp), __ldg(stp + 1));
      }
      const u32 longest = __reduce_max_sync(0xffffffffu, rA.y - rA.x);
      const u32 far = ~fb;  // at distance 32 from the lane's beta string
      for (u32 i = 0; i < longest; i += 4) {
        u32 k[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const u32 p = rA.x + i + (u32)u;
          k[u] = p < rA.y ? __ldg(halfA + p) : far;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const u32 d = (u32)__popc(k[u] ^ fb);
          if (d == 2u || d == 4u) emit_hit(c, h, (rA.x + i + (u32)u) | kHitA | kHitOwn);
        }
      }
    }
    if (active) {
      const u32 first = part * runs_per_part;
      if (h.over) {
        c.my_runs[first] = HitRun{0u, kOverflow};
        h.nrun = first + 1u;
      } else if (h.fill) {
        c.my_runs[h.nrun] = HitRun{h.base, h.fill};
        ++h.nrun;
      }
      if (parts == 1u) {
        run_cnt[sid] = h.nrun;
      } else {
        // every part leaves its unused records empty and reports all of them (the same value from every part)
        for (u32 r = h.nrun; r < h.end; ++r) c.my_runs[r] = HitRun{0u, 0u};
        run_cnt[sid] = (u32)run_stride;
      }
    }
    __syncwarp();
  }
}

// ---- host side --------------------------------------------------------------------------------------------------------
struct BlockLayout {
  u32 log2_buckets;
  long long bcnt, bbase, slots, tiles, total;
};

BlockLayout block_layout(long long n) {
  BlockLayout l;
  l.log2_buckets = sample_log2_buckets(n);
  auto up = [](long long v) { return (v + 255) & ~255LL; };
  long long o = 0;
  l.bcnt = o;
  o = up(o + 4LL * (1LL << l.log2_buckets));
  l.bbase = o;
  o = up(o + 4LL * (1LL << l.log2_buckets));
  l.slots = o;
  o = up(o + 4 * n);
  l.tiles = o;
  o = up(o + 8 * (n / 4 + 64));
  l.total = o;
  return l;
}

long long block_scratch_bytes(long long n) { return block_layout(n).total; }
int block_run_stride() { return 8; }

// grouping pass + block kernel over all n samples; the samples it leaves to the per-sample kernel are the last
// ctr->n_single entries of the slot array (returned through *slots_out)
int launch_eloc_block(const u64 *bra, long long n, const GroupView &gv, char *block_ws, ElocCounters *ctr, HitRun *runs, u32 *run_cnt,
                      int run_stride, u32 *hits, u32 *self_pos, u32 hit_cap, const ExcGeom &g, const u32 **slots_out, cudaStream_t st) {
  const BlockLayout l = block_layout(n);
  u32 *bcnt = reinterpret_cast<u32 *>(block_ws + l.bcnt), *bbase = reinterpret_cast<u32 *>(block_ws + l.bbase);
  u32 *slots = reinterpret_cast<u32 *>(block_ws + l.slots);
  uint2 *tiles = reinterpret_cast<uint2 *>(block_ws + l.tiles);
  *slots_out = slots;
  const u32 nbuckets = 1u << l.log2_buckets, shift = 32u - l.log2_buckets;
  if (cudaMemsetAsync(bcnt, 0, 4 * (size_t)nbuckets, st) != cudaSuccess) return check_launch("eloc block memset");
  const unsigned sb = (unsigned)((n + 255) / 256);
  sample_count_kernel<<<sb, 256, 0, st>>>(bra, n, shift, bcnt);
  int min_group = eloc_tuning().block_min_group;
  if (min_group < 4) min_group = 4;  // tiles[] holds n / 4 + 64 entries
  sample_alloc_kernel<<<(nbuckets + 255) / 256, 256, 0, st>>>(bcnt, bbase, nbuckets, ctr, tiles, (u32)min_group, (u32)n);
  sample_scatter_kernel<<<sb, 256, 0, st>>>(bra, n, shift, bcnt, bbase, slots);
  count_launch(3);
  if (int rc = check_launch("eloc grouping pass")) return rc;
  BlockGeom bg;
  bg.sorb = g.sorb;
  bg.noA = g.noA;
  bg.noB = g.noB;
  bg.nvB = g.nvB;
  // idle lanes and pads must never pass a test, against real strings (noA bits) and against each other
  bg.never = g.noA > 4 ? 0xffffffffu : 0x0000ffffu;  // distance >= 16 - noA >= 12 from every key, >= 16 from the pad
  bg.pad = g.noA > 4 ? 0u : 0xffffffffu;             // distance noA resp. 32 - noA from every alpha string
  bg.pow2 = g.noA > 4;
  const size_t smem = sizeof(BlockWarpSmem) * kBlkWarps;
  long long ctas = (n + 32 * kBlkWarps - 1) / (32 * kBlkWarps);
  const long long cap = 148LL * (long long)((227 * 1024) / (sizeof(BlockWarpSmem) * kBlkWarps + 1024));
  if (ctas > cap) ctas = cap;
  // fewer than ~2 tiles (of ~26 samples) per resident warp: split the tiles (the knob block_parts forces 1, 2 or 4 parts)
  const int forced = eloc_tuning().block_parts;
  const bool split = forced ? forced > 1 : n < 2 * 26 * ctas * kBlkWarps;
  auto kern = split ? eloc_block_kernel<true> : eloc_block_kernel<false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return check_launch("eloc_block_kernel smem opt-in");
  kern<<<(unsigned)ctas, kBlkWarps * 32, smem, st>>>(bra, slots, tiles, ctr, gv, runs, run_cnt, run_stride, hits, self_pos, hit_cap, bg,
                                                     (u32)(forced == 3 ? 2 : forced));
  count_launch();
  return check_launch("eloc_block_kernel");
}

}  // namespace pynqs
