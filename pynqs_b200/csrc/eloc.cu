// eloc.cu -- sample-space local energy without materialising comb / Hmat.
//
// Replaces what the reference does in three native calls plus ~6 torch passes over [n, M]
// (vmc/energy/eloc.py:326-397: get_comb_hij_fused -> WavefunctionLUT.lookup -> scatter ->
// divide -> multiply -> sum).  Nothing of size [n, M] ever reaches HBM.  Two kernels:
//
//   eloc_filter_kernel   one CTA per sample.  Builds the per-sample excitation tables
//       (tables.cuh), enumerates every connected determinant and probes the FIRST bucket of its
//       region in the string-grouped index (lut.cuh).  A row dies here unless one of the four
//       32-bit tags matches or the bucket has overflowed -- >99 % of the rows for a sparse table.
//       The loop has no calls and no cold code, so it needs few registers and the latency of the
//       probes is hidden by occupancy plus kElocRows independent probes per lane.  Survivors
//       ("candidates") are compacted with ballot + popcount prefix into per-warp queues and
//       written to a global candidate buffer in a deterministic order.
//   eloc_eval_kernel     one warp per sample.  For each candidate: decode the excitation, finish
//       the probe sequence, compare with the key table, and on a hit add (psi(x')/psi(x)) * <x|H|x'>
//       with <x|H|x'> computed exactly as the fused operator does (same order of additions).
//
// Determinism: fixed row -> lane mapping, candidates in a fixed order, shuffle tree; no floating
// point atomics.  Per element the arithmetic is the reference's (psi'/psi0 first, then * H); only
// the order of the final sum differs from torch's reduction (tolerance 1e-12 rel).
#include "lut.cuh"
#include "tables.cuh"

namespace pynqs {

constexpr int kElocThreads = 256;         // filter kernel: 8 warps per sample
constexpr int kElocWarps = kElocThreads / 32;
constexpr int kElocRows = 4;              // independent probes in flight per lane
constexpr int kCandPerWarp = 256;         // per-warp candidate queue (overflow -> the sample is re-done in full)
constexpr int kEvalThreads = 128;         // eval kernel: 4 samples per CTA
constexpr u32 kCandOverflow = 0x80000000u;

struct Cplx {
  double re, im;
};

// numpy / c10 complex division (torch/headeronly/util/complex.h operator/=)
__device__ __forceinline__ Cplx cdiv(Cplx x, Cplx y) {
  const double a = x.re, b = x.im, c = y.re, d = y.im;
  const double ac = fabs(c), ad = fabs(d);
  Cplx r;
  if (ac >= ad) {
    if (ac == 0.0 && ad == 0.0) {
      r.re = a / ac;
      r.im = b / ad;
    } else {
      const double rat = d / c, scl = 1.0 / (c + d * rat);
      r.re = (a + b * rat) * scl;
      r.im = (b - a * rat) * scl;
    }
  } else {
    const double rat = c / d, scl = 1.0 / (d + c * rat);
    r.re = (a * rat + b) * scl;
    r.im = (b * rat - a) * scl;
  }
  return r;
}

template <bool CPLX>
__device__ __forceinline__ Cplx load_psi(const double *__restrict__ psi, long long id) {
  Cplx v;
  if (CPLX) {
    const double2 t = __ldg(reinterpret_cast<const double2 *>(psi) + id);
    v.re = t.x;
    v.im = t.y;
  } else {
    v.re = __ldg(psi + id);
    v.im = 0.0;
  }
  return v;
}

// ratio psi'/psi0 times a real H, accumulated
template <bool CPLX>
__device__ __forceinline__ void accumulate(Cplx &acc, Cplx pm, Cplx p0, double h) {
  if (CPLX) {
    const Cplx q = cdiv(pm, p0);
    acc.re += q.re * h;
    acc.im += q.im * h;
  } else {
    acc.re += (pm.re / p0.re) * h;
  }
}

// ---- per-sample tables of the filter kernel: two dense u64 arrays indexed like tables.cuh ------------
// aux[t]: SA -> hash of the excited ALPHA string alpha(x) ^ m (tag / bucket inside a beta-grouped region)
//         SB -> probe descriptor of the region of the excited BETA string beta(x) ^ m (kNoRegion: no key
//               of the table has that beta string, so all noA*nvA alpha-beta doubles on top of it miss)
//         pair tables -> unused
// msk[t]: the entry's two orbitals as an excitation mask (L = 1) or packed o0 | o1 << 8 (L > 1); only
//         the classes that must hash a freshly excited string read it (beta singles, same-spin doubles).
// 8-byte entries: the 32 lanes of a warp read 32 consecutive entries = 256 contiguous bytes, no bank conflicts.
template <int L>
__device__ __forceinline__ u64 msk_make(u32 o0, u32 o1) {
  return L == 1 ? ((1ull << o0) | (1ull << o1)) : (u64)(o0 | (o1 << 8));
}
template <int L>
__device__ __forceinline__ Onv<L> msk_apply(const Onv<L> &x, u64 m) {
  Onv<L> y = x;
  if (L == 1) {
    y.w[0] ^= m;
  } else {
    flip_bit<L>(y, (int)(m & 0xffu));
    flip_bit<L>(y, (int)((m >> 8) & 0xffu));
  }
  return y;
}

// probe descriptor: lo = first bucket of the region in the pool, hi = 32 - log2(buckets) (so
// bucket = hash_hi >> hi and mask = 0xffffffff >> hi); kNoRegion: the string is not in the table
__device__ __forceinline__ u64 pack_probe_desc(u64 region_desc) {
  if (region_desc == kNoRegion) return kNoRegion;
  return (region_desc & 0xffffffffull) | ((u64)(32u - (u32)(region_desc >> 32)) << 32);
}
__device__ __forceinline__ u32 shr_clamp(u32 v, u32 s) {  // PTX shr: shift amounts >= 32 give 0
  u32 r;
  asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
  return r;
}

// per (sample, split, warp) candidate run in the global buffer
struct CandRun {
  u32 off, cnt;  // cnt & kCandOverflow: queue or buffer overflowed -> evaluate every row of the split
};

// One chunk of 32 * ROWS consecutive rows of one excitation class (CLS 0/1 single a/b, 2/3 double
// aa/bb, 4 double ab; excitation index r = row), lane l owning rows c + l + 32 j: issue the first
// bucket load of ROWS probes per lane, then test the four tags of each bucket.  FULL: every row of
// the chunk is inside [lo, hi).
template <int L, int CLS, int ROWS, bool FULL>
__device__ __forceinline__ void eloc_filter_chunk(const Onv<L> &x, const u64 *aux, const u64 *msk, const TableOffsets &to,
                                                  const ExcGeom &g, u64 own, const TagBucket *__restrict__ pool, u32 *queue,
                                                  u32 &qn, int c, int hi) {
  const int lane = threadIdx.x & 31;
  uint4 t[ROWS];
  u32 tag[ROWS];  // 0: no probe for this row
#pragma unroll
  for (int j = 0; j < ROWS; ++j) {
    const int r = c + lane + 32 * j;
    u64 desc = kNoRegion, h2 = 0;
    if (FULL || r < hi) {
      if (CLS == 0) {
        desc = own;
        h2 = aux[to.sa + r];
      } else if (CLS == 1) {
        desc = own;
        h2 = hash_beta<L>(msk_apply<L>(x, msk[to.sb + (r - g.d0)]));
      } else if (CLS == 2) {
        const u64 m1 = msk[to.hpa + (int)((u32)r - fdiv((u32)r, g.by_noAA) * g.noAA)];  // global r (quirk Q1)
        const u64 m2 = msk[to.ppa + (int)fdiv((u32)(r - g.d1), g.by_noAA)];
        desc = own;
        h2 = hash_alpha<L>(msk_apply<L>(msk_apply<L>(x, m1), m2));
      } else if (CLS == 3) {
        const u64 m1 = msk[to.hpb + (int)((u32)r - fdiv((u32)r, g.by_noBB) * g.noBB)];
        const u64 m2 = msk[to.ppb + (int)fdiv((u32)(r - g.d2), g.by_noBB)];
        desc = own;
        h2 = hash_beta<L>(msk_apply<L>(msk_apply<L>(x, m1), m2));
      } else {
        const u32 q = (u32)(r - g.d3);
        const u32 jb = fdiv(q, g.by_sA);
        h2 = aux[to.sa + (int)(q - jb * g.sA)];  // hash of the excited alpha string
        desc = aux[to.sb + (int)jb];             // probe descriptor of the excited beta string's region
      }
    }
    const u32 sh32 = (u32)(desc >> 32);
    tag[j] = 0;
    if (sh32 != ~0u) {
      tag[j] = hash_tag(h2);
      t[j] = __ldg(pool + (u32)desc + shr_clamp((u32)(h2 >> 32), sh32));
    }
  }
  bool cand[ROWS];
  bool any = false;
#pragma unroll
  for (int j = 0; j < ROWS; ++j) {
    cand[j] = tag[j] && probe_residue(t[j], tag[j]) == 0u;
    any |= cand[j];
  }
  if (__any_sync(0xffffffffu, any)) {
    const u32 lt = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      const u32 m = __ballot_sync(0xffffffffu, cand[j]);
      const u32 slot = qn + __popc(m & lt);
      if (cand[j] && slot < (u32)kCandPerWarp) queue[slot] = (u32)(c + lane + 32 * j);
      qn += __popc(m);  // may exceed the capacity: recorded as overflow at the end
    }
  }
}

template <int L, int CLS, int ROWS>
__device__ __forceinline__ void eloc_filter_class(const Onv<L> &x, const u64 *aux, const u64 *msk, const TableOffsets &to,
                                                  const ExcGeom &g, u64 own, const TagBucket *__restrict__ pool, u32 *queue,
                                                  u32 &qn, int lo, int hi) {
  if (CLS != 4 && own == kNoRegion) return;  // the sample's own string is not in the table: nothing of this class is
  constexpr int kChunk = 32 * ROWS;
  for (int c = lo + (int)(threadIdx.x >> 5) * kChunk; c < hi; c += kElocWarps * kChunk) {
    if (c + kChunk <= hi) eloc_filter_chunk<L, CLS, ROWS, true>(x, aux, msk, to, g, own, pool, queue, qn, c, hi);
    else eloc_filter_chunk<L, CLS, ROWS, false>(x, aux, msk, to, g, own, pool, queue, qn, c, hi);
  }
}

// Same-spin doubles (BETA = false: alpha-alpha rows [d1, d2), own beta region; true: beta-beta rows
// [d2, d3), own alpha region).  Row r = base + ab * nH + k uses particle pair ab and hole pair
// ij = r % nH = (k + base) % nH (the reference takes the GLOBAL index modulo nH, quirk Q1), i.e. the
// hole pairs of one particle pair are cyclically shifted by base % nH.  Work unit = (particle pair,
// block of 32 hole pairs): lane l owns hole pair ij = 32 blk + l, so there is no division per row;
// the particle-pair mask is warp-uniform.  Two units are in flight per iteration.
template <int L, bool BETA>
__device__ __forceinline__ void eloc_filter_ss(const Onv<L> &x, const u64 *msk, const TableOffsets &to, const ExcGeom &g, u64 own,
                                               const TagBucket *__restrict__ pool, u32 *queue, u32 &qn, int lo, int hi) {
  if (lo >= hi || own == kNoRegion) return;
  constexpr int U = 2;
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 nH = (u32)(BETA ? g.noBB : g.noAA), nP = (u32)(BETA ? g.nvBB : g.nvAA);
  const u32 base = (u32)(BETA ? g.d2 : g.d1);
  const u64 *hp = msk + (BETA ? to.hpb : to.hpa), *pp = msk + (BETA ? to.ppb : to.ppa);
  const u32 shift = base - fdiv(base, BETA ? g.by_noBB : g.by_noAA) * nH;  // base % nH
  const u32 nblk = (nH + 31u) >> 5;
  const u32 roff = (u32)own, sh32 = (u32)(own >> 32);
  // unit u = ab * nblk + blk; this warp takes units warp, warp + 8, ... -- (ab, blk) advanced incrementally
  const u32 step_ab = (u32)kElocWarps / nblk, step_blk = (u32)kElocWarps - step_ab * nblk;
  u32 ab = warp / nblk, blk = warp - ab * nblk;
  while (ab < nP) {
    uint4 t[U];
    u32 tag[U], row[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      tag[i] = 0;
      row[i] = 0;
      const u32 ij = (blk << 5) + lane;
      if (ab < nP && ij < nH) {
        const u32 k = ij >= shift ? ij - shift : ij + nH - shift;
        const u32 r = base + ab * nH + k;
        if (r >= (u32)lo && r < (u32)hi) {
          const Onv<L> y = msk_apply<L>(msk_apply<L>(x, hp[ij]), pp[ab]);
          const u64 h2 = BETA ? hash_beta<L>(y) : hash_alpha<L>(y);
          row[i] = r;
          tag[i] = hash_tag(h2);
          const u32 bucket = roff + shr_clamp((u32)(h2 >> 32), sh32);
          t[i] = __ldg(pool + bucket);
        }
      }
      ab += step_ab;
      blk += step_blk;
      if (blk >= nblk) {
        blk -= nblk;
        ++ab;
      }
    }
    u32 res[U];
    bool any = false;
#pragma unroll
    for (int i = 0; i < U; ++i) {
      res[i] = tag[i] ? probe_residue(t[i], tag[i]) : 1u;
      any |= res[i] == 0u;
    }
    if (__any_sync(0xffffffffu, any)) {
      const u32 lt = (1u << lane) - 1u;
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const u32 m = __ballot_sync(0xffffffffu, res[i] == 0u);
        const u32 slot = qn + __popc(m & lt);
        if (res[i] == 0u && slot < (u32)kCandPerWarp) queue[slot] = row[i];
        qn += __popc(m);
      }
    }
  }
}

// Alpha-beta doubles, whole runs: rows r = d3 + jb * sA + ia for jb in [jb_begin, jb_end), all ia.
// The warp walks the beta excitations jb assigned to it.  The region of the excited beta string is
// warp-uniform (one broadcast shared load; when no key has that string the whole run of sA rows is
// skipped), and lane l always handles the alpha excitations ia = l + 32 j, so the hashes of ITS
// excited alpha strings stay in registers for the whole sample: a row costs a shift, a 64-bit
// multiply-add for the address and the 16-byte gather of its four tags, then a branch-free test
// (probe_residue) -- no division, no per-row shared-memory traffic.  Two runs are in flight.
// Requires sA <= 32 * RA (RA <= 4); larger alpha tables use the generic row loop.
template <int L, int RA>
__device__ __forceinline__ void eloc_filter_ab_runs(const u64 *aux_sa, const u64 *aux_sb, const ExcGeom &g,
                                                    const TagBucket *__restrict__ pool, u32 *queue, u32 &qn, u32 jb_begin, u32 jb_end) {
  constexpr int JB = 2;
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 sA = (u32)g.sA;
  u32 tag[RA], hhi[RA];
#pragma unroll
  for (int j = 0; j < RA; ++j) {
    // slots beyond sA repeat the last valid one: harmless duplicate probes, masked out when queued
    const u32 ia = min(lane + 32u * j, sA - 1u);
    const uint2 h = *reinterpret_cast<const uint2 *>(aux_sa + ia);
    tag[j] = h.x | 3u;
    hhi[j] = h.y;
  }
  for (u32 jb0 = jb_begin + warp; jb0 < jb_end; jb0 += JB * kElocWarps) {
    uint4 t[JB][RA];
    bool live[JB];
#pragma unroll
    for (int b = 0; b < JB; ++b) {
      const u32 jb = jb0 + b * kElocWarps;
      uint2 d = make_uint2(0u, ~0u);
      if (jb < jb_end) d = *reinterpret_cast<const uint2 *>(aux_sb + jb);  // (first bucket, 32 - log2 buckets), warp-uniform
      live[b] = d.y != ~0u;
      if (live[b]) {
#pragma unroll
        for (int j = 0; j < RA; ++j) {
          const u32 bucket = d.x + shr_clamp(hhi[j], d.y);  // 32-bit index arithmetic, one 64-bit address per load
          t[b][j] = __ldg(pool + bucket);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < JB; ++b) {
      if (!live[b]) continue;  // warp-uniform
      u32 res[RA];
      u32 worst = 1u;
#pragma unroll
      for (int j = 0; j < RA; ++j) {
        res[j] = probe_residue(t[b][j], tag[j]);
        worst = min(worst, res[j]);
      }
      if (__any_sync(0xffffffffu, worst == 0u)) {
        const u32 lt = (1u << lane) - 1u;
        const u32 jb = jb0 + b * kElocWarps;
#pragma unroll
        for (int j = 0; j < RA; ++j) {
          const bool c = res[j] == 0u && lane + 32u * j < sA;
          const u32 m = __ballot_sync(0xffffffffu, c);
          const u32 slot = qn + __popc(m & lt);
          if (c && slot < (u32)kCandPerWarp) queue[slot] = (u32)g.d3 + jb * sA + lane + 32 * j;
          qn += __popc(m);  // may exceed the capacity: recorded as overflow at the end
        }
      }
    }
  }
}

// alpha-beta class of a split [lo, hi): whole runs through the register-resident fast path, the
// partial runs at the split boundaries (and alpha tables beyond 128 entries) through the row loop
template <int L>
__device__ __forceinline__ void eloc_filter_ab(const Onv<L> &x, const u64 *aux, const u64 *msk, const TableOffsets &to,
                                               const ExcGeom &g, const TagBucket *__restrict__ pool, u32 *queue, u32 &qn, int lo,
                                               int hi) {
  if (lo >= hi) return;
  const u32 sA = (u32)g.sA;
  if (sA > 128u) {
    eloc_filter_class<L, 4, kElocRows>(x, aux, msk, to, g, kNoRegion, pool, queue, qn, lo, hi);
    return;
  }
  const u32 q_lo = (u32)(lo - g.d3), q_hi = (u32)(hi - g.d3);
  const u32 jb_begin = fdiv(q_lo + sA - 1u, g.by_sA), jb_end = fdiv(q_hi, g.by_sA);  // whole runs [jb_begin, jb_end)
  if (jb_begin >= jb_end) {  // the split holds no whole run
    eloc_filter_class<L, 4, kElocRows>(x, aux, msk, to, g, kNoRegion, pool, queue, qn, lo, hi);
    return;
  }
  const int head_hi = g.d3 + (int)(jb_begin * sA);
  if (lo < head_hi) eloc_filter_class<L, 4, kElocRows>(x, aux, msk, to, g, kNoRegion, pool, queue, qn, lo, head_hi);
  const u64 *aux_sa = aux + to.sa, *aux_sb = aux + to.sb;
  if (sA <= 32u) eloc_filter_ab_runs<L, 1>(aux_sa, aux_sb, g, pool, queue, qn, jb_begin, jb_end);
  else if (sA <= 64u) eloc_filter_ab_runs<L, 2>(aux_sa, aux_sb, g, pool, queue, qn, jb_begin, jb_end);
  else if (sA <= 96u) eloc_filter_ab_runs<L, 3>(aux_sa, aux_sb, g, pool, queue, qn, jb_begin, jb_end);
  else eloc_filter_ab_runs<L, 4>(aux_sa, aux_sb, g, pool, queue, qn, jb_begin, jb_end);
  const int tail_lo = g.d3 + (int)(jb_end * sA);
  if (tail_lo < hi) eloc_filter_class<L, 4, kElocRows>(x, aux, msk, to, g, kNoRegion, pool, queue, qn, tail_lo, hi);
}

template <int L>
__global__ void __launch_bounds__(kElocThreads, 4)
eloc_filter_kernel(const u64 *__restrict__ bra, long long n, IndexView iv, CandRun *__restrict__ runs, u32 *__restrict__ cand,
                   u32 *cursor, u32 cand_cap, int splits, ExcGeom g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const TableOffsets to = table_offsets(g);
  OrbLists &lists = *reinterpret_cast<OrbLists *>(smem_raw);
  u64 *aux = reinterpret_cast<u64 *>(smem_raw + sizeof(OrbLists));
  u64 *msk = aux + to.total;
  u32 *queues = reinterpret_cast<u32 *>(msk + to.total);
  __shared__ u64 s_own[2];  // probe descriptors of the regions of the sample's own beta / alpha string

  const long long s = splits == 1 ? (long long)blockIdx.x : (long long)(blockIdx.x / (unsigned)splits);
  const int split = splits == 1 ? 0 : (int)(blockIdx.x - (unsigned)s * (unsigned)splits);
  if (s >= n) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  CandRun *my_run = runs + (s * splits + split) * kElocWarps + warp;
  if (iv.hdr->has_dup) {  // duplicate keys: the eval kernel redoes every row with the reference's classic search
    if (lane == 0) {
      const CandRun run = {0u, kCandOverflow};
      *my_run = run;
    }
    return;
  }
  const Onv<L> x = load_onv<L>(bra + s * L);
  if (threadIdx.x < 32) build_lists<L>(x, g.sorb, g.noA, g.noB, lists, threadIdx.x);
  // the two directory probes of the sample itself run beside the table build (different warps)
  if (threadIdx.x == kElocThreads - 1) s_own[0] = pack_probe_desc(dir_find(iv.dir[0], iv.log2_dir, hash_beta<L>(x)));
  if (threadIdx.x == kElocThreads - 33) s_own[1] = pack_probe_desc(dir_find(iv.dir[1], iv.log2_dir, hash_alpha<L>(x)));
  __syncthreads();
  for_each_table_entry(g, lists, to, [&](int t, int kind, u32 e0, u32 e1) {
    const u64 m = msk_make<L>(e0 & 0xffu, e1 & 0xffu);
    u64 a = 0;
    if (kind == 0) a = hash_alpha<L>(msk_apply<L>(x, m));
    else if (kind == 1) a = pack_probe_desc(dir_find(iv.dir[0], iv.log2_dir, hash_beta<L>(msk_apply<L>(x, m))));
    aux[t] = a;
    msk[t] = m;
  });
  __syncthreads();

  const int chunk = splits == 1 ? g.nsd : (int)(((unsigned)g.nsd + (unsigned)splits - 1u) / (unsigned)splits);
  const int r_begin = split * chunk;
  const int r_end = min(g.nsd, r_begin + chunk);
  u32 *queue = queues + warp * kCandPerWarp;
  u32 qn = 0;
  const u64 own_b = s_own[0], own_a = s_own[1];
  auto lo = [&](int v) { return v > r_begin ? v : r_begin; };
  auto hi = [&](int v) { return v < r_end ? v : r_end; };
  eloc_filter_ab<L>(x, aux, msk, to, g, iv.pool, queue, qn, lo(g.d3), hi(g.nsd));
  eloc_filter_ss<L, false>(x, msk, to, g, own_b, iv.pool, queue, qn, lo(g.d1), hi(g.d2));
  eloc_filter_ss<L, true>(x, msk, to, g, own_a, iv.pool, queue, qn, lo(g.d2), hi(g.d3));
  eloc_filter_class<L, 0, 1>(x, aux, msk, to, g, own_b, iv.pool, queue, qn, lo(0), hi(g.d0));
  eloc_filter_class<L, 1, 1>(x, aux, msk, to, g, own_a, iv.pool, queue, qn, lo(g.d0), hi(g.d1));

  // publish this warp's candidates (no CTA barrier: every warp owns its run record)
  __syncwarp();
  u32 off = 0;
  bool over = qn > (u32)kCandPerWarp;
  if (lane == 0 && !over && qn) {
    off = atomicAdd(cursor, qn);
    if (off > cand_cap || qn > cand_cap - off) over = true;  // buffer exhausted: full evaluation for this run
  }
  off = __shfl_sync(0xffffffffu, off, 0);
  over = __shfl_sync(0xffffffffu, (int)over, 0) != 0;
  if (lane == 0) {
    const CandRun run = {off, over ? kCandOverflow : qn};
    *my_run = run;
  }
  if (!over)
    for (u32 e = lane; e < qn; e += 32) cand[off + e] = queue[e];
}

// ---- evaluation of the candidates ------------------------------------------------------------------------
// full lookup of determinant y = x after excitation r; own_b / own_a: probe descriptors of the sample's
// own beta / alpha regions.  r < 0: y is the sample itself.
template <int L>
__device__ __forceinline__ long long eloc_lookup(const Onv<L> &y, int r, const ExcGeom &g, u64 own_b, u64 own_a, const IndexView &iv,
                                                 bool dup, const u64 *__restrict__ key, long long N) {
  if (dup) return classic_search<L>(key, N, y);
  u64 desc, h2;
  if (r >= g.d3) {  // alpha-beta double: region of the excited beta string
    desc = pack_probe_desc(dir_find(iv.dir[0], iv.log2_dir, hash_beta<L>(y)));
    h2 = hash_alpha<L>(y);
  } else if (r < g.d0 || (r >= g.d1 && r < g.d2)) {  // the sample itself / alpha single / alpha-alpha double: own beta string
    desc = own_b;
    h2 = hash_alpha<L>(y);
  } else {  // beta single / beta-beta double: own alpha string
    desc = own_a;
    h2 = hash_beta<L>(y);
  }
  if (desc == kNoRegion) return -1;
  const u32 roff = (u32)desc, sh32 = (u32)(desc >> 32), msk = shr_clamp(~0u, sh32), tag = hash_tag(h2);
  u32 b = shr_clamp((u32)(h2 >> 32), sh32);
  for (u32 probe = 0;; ++probe) {
    const uint4 tj = __ldg(iv.pool + roff + b);
    if (tags_match(tj, tag)) {
      const uint4 ids = __ldg(iv.rows + roff + b);
      const u32 tg[4] = {tj.x, tj.y, tj.z, tj.w | 1u}, id[4] = {ids.x, ids.y, ids.z, ids.w};
#pragma unroll
      for (int sl = 0; sl < 4; ++sl)
        if (tg[sl] == tag && eq_onv<L>(load_onv<L>(key + (long long)id[sl] * L), y)) return (long long)id[sl];
    }
    if (!bucket_overflowed(tj) || probe >= msk) return -1;
    b = (b + 1) & msk;
  }
}

template <int L, bool CPLX>
__device__ __forceinline__ void eloc_eval_row(const Onv<L> &x, const OrbLists &lists, const ExcGeom &g, u64 own_b, u64 own_a,
                                              Cplx p0, const IndexView &iv, bool dup, const u64 *__restrict__ key, long long N,
                                              const double *__restrict__ psi, const double *__restrict__ h1e,
                                              const double *__restrict__ h2e, int r, Cplx &acc) {
  const Exc e = decode_exc(g, lists, r);
  const long long id = eloc_lookup<L>(apply_exc<L>(x, e), r, g, own_b, own_a, iv, dup, key, N);
  if (id >= 0) accumulate<CPLX>(acc, load_psi<CPLX>(psi, id), p0, exc_element<L, double>(x, e, h1e, h2e, g.sorb));
}

template <int L, bool CPLX>
__global__ void __launch_bounds__(kEvalThreads)
eloc_eval_kernel(const u64 *__restrict__ bra, long long n, const double *__restrict__ h1e, const double *__restrict__ h2e,
                 const u64 *__restrict__ key, const double *__restrict__ psi, long long N, IndexView iv,
                 const CandRun *__restrict__ runs, const u32 *__restrict__ cand, const double *__restrict__ hii,
                 double *__restrict__ eloc, double *__restrict__ psi0_out, int splits, ExcGeom g) {
  __shared__ OrbLists s_lists[kEvalThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long s = (long long)blockIdx.x * (kEvalThreads / 32) + warp;
  if (s >= n) return;
  OrbLists &lists = s_lists[warp];
  const Onv<L> x = load_onv<L>(bra + s * L);
  build_lists<L>(x, g.sorb, g.noA, g.noB, lists, lane);
  __syncwarp();
  const bool dup = iv.hdr->has_dup != 0;
  // the sample itself: its two regions and psi0 (every lane computes the same values)
  const u64 own_b = dup ? kNoRegion : pack_probe_desc(dir_find(iv.dir[0], iv.log2_dir, hash_beta<L>(x)));
  const u64 own_a = dup ? kNoRegion : pack_probe_desc(dir_find(iv.dir[1], iv.log2_dir, hash_alpha<L>(x)));
  const long long id0 = eloc_lookup<L>(x, -1, g, own_b, own_a, iv, dup, key, N);
  Cplx p0 = {0.0, 0.0};
  if (id0 >= 0) p0 = load_psi<CPLX>(psi, id0);
  Cplx acc = {0.0, 0.0};
  if (lane == 0) accumulate<CPLX>(acc, p0, p0, hii[s]);  // row 0: (psi0/psi0) * H_xx
  const int chunk = (g.nsd + splits - 1) / splits;
  for (int k = 0; k < splits; ++k) {
    const int r_begin = k * chunk, r_end = min(g.nsd, r_begin + chunk);
    bool redo = false;
    for (int w = 0; w < kElocWarps; ++w) redo |= (runs[(s * splits + k) * kElocWarps + w].cnt & kCandOverflow) != 0;
    if (redo) {  // some queue overflowed (or duplicate keys): every row of the split, in row order
      for (int r = r_begin + lane; r < r_end; r += 32)
        eloc_eval_row<L, CPLX>(x, lists, g, own_b, own_a, p0, iv, dup, key, N, psi, h1e, h2e, r, acc);
    } else {
      // the split's candidates = its kElocWarps runs back to back; lane w < kElocWarps holds run w and
      // an exclusive prefix of the counts, so the whole list is walked 32 candidates at a time
      CandRun mine = {0u, 0u};
      if (lane < kElocWarps) mine = runs[(s * splits + k) * kElocWarps + lane];
      u32 incl = mine.cnt;
#pragma unroll
      for (int o = 1; o < kElocWarps; o <<= 1) {
        const u32 up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      const u32 total = __shfl_sync(0xffffffffu, incl, kElocWarps - 1);
      for (u32 e0 = 0; e0 < total; e0 += 32) {
        const u32 e = e0 + lane;
        // run that holds candidate e: the first w with incl[w] > e
        u32 w = 0, start = 0, roff = 0;
#pragma unroll
        for (int q = 0; q < kElocWarps; ++q) {
          const u32 iq = __shfl_sync(0xffffffffu, incl, q), oq = __shfl_sync(0xffffffffu, mine.off, q);
          const u32 cq = __shfl_sync(0xffffffffu, mine.cnt, q);
          if (e >= iq - cq && e < iq) { w = (u32)q; start = iq - cq; roff = oq; }
        }
        (void)w;
        if (e < total)
          eloc_eval_row<L, CPLX>(x, lists, g, own_b, own_a, p0, iv, dup, key, N, psi, h1e, h2e, (int)cand[roff + (e - start)], acc);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc.re += __shfl_xor_sync(0xffffffffu, acc.re, o);
    if (CPLX) acc.im += __shfl_xor_sync(0xffffffffu, acc.im, o);
  }
  if (lane == 0) {
    if (CPLX) {
      eloc[2 * s] = acc.re;
      eloc[2 * s + 1] = acc.im;
      psi0_out[2 * s] = p0.re;
      psi0_out[2 * s + 1] = p0.im;
    } else {
      eloc[s] = acc.re;
      psi0_out[s] = p0.re;
    }
  }
}

// ---- host side --------------------------------------------------------------------------------------------
constexpr long long kElocMaxCtas = 1 << 18;  // (samples x splits) per filter launch
constexpr int kRowsPerSplit = 8192;         // a CTA filters at most this many rows (1024 per warp, queue of 256)

// splits per sample: enough CTAs to fill the GPU when n is small, and never more than kRowsPerSplit
// rows per CTA so that the per-warp candidate queues only overflow at hit rates above ~25 %
int eloc_splits(long long n, int nsd) {
  if (n <= 0) return 1;
  long long want = (148LL * 8 + n - 1) / n;
  const long long cap = (nsd + 2047) / 2048 > 1 ? (nsd + 2047) / 2048 : 1;
  if (want > cap) want = cap;
  const long long need = (nsd + kRowsPerSplit - 1) / kRowsPerSplit;
  if (want < need) want = need;
  if (want < 1) want = 1;
  return (int)want;
}

static long long eloc_batch(int nsd) {
  const long long per_sample = (nsd + kRowsPerSplit - 1) / kRowsPerSplit > 1 ? (nsd + kRowsPerSplit - 1) / kRowsPerSplit : 1;
  long long b = kElocMaxCtas / per_sample;
  return b < 1024 ? 1024 : b;
}

struct ElocScratch {
  long long hii, runs, cursor, cand, total;
  long long cand_cap, batch;
  int splits;  // same for every batch of the call (sized for the first, largest one)
};

static ElocScratch eloc_scratch_layout(long long n, int nsd) {
  ElocScratch l;
  l.batch = eloc_batch(nsd);
  const long long nb = n < l.batch ? n : l.batch;
  const int splits = eloc_splits(nb, nsd);
  l.splits = splits;
  l.hii = 0;
  l.runs = (l.hii + 8 * n + 15) / 16 * 16;
  l.cursor = l.runs + (long long)sizeof(CandRun) * nb * splits * kElocWarps;
  l.cand = l.cursor + 256;
  // candidates per sample the global buffer can take before runs fall back to full evaluation:
  // 160 (hits + overflowed buckets of a sparse table) or 1/32 of the rows, whichever is larger
  const long long per_sample = nsd / 32 > 160 ? nsd / 32 : 160;
  l.cand_cap = nb * per_sample + 65536;
  if (l.cand_cap > 0x7fffffffLL) l.cand_cap = 0x7fffffffLL;
  l.total = l.cand + 4 * l.cand_cap + 256;
  return l;
}

long long eloc_scratch_bytes(long long n, int nsd, int) { return eloc_scratch_layout(n, nsd).total; }

int launch_diag_f64(const u64 *bra, const double *h1e, const double *h2e, double *out, long long n, long long stride, int L,
                    int sorb, int nele, cudaStream_t st);

template <int L, bool CPLX>
static int launch_eloc_LC(const u64 *bra, long long n, const double *h1e, const double *h2e, const u64 *key, const double *psi,
                          long long N, const IndexView &iv, char *scratch, const ElocScratch &lay, double *eloc, double *psi0,
                          const ExcGeom &g, cudaStream_t st) {
  double *hii = reinterpret_cast<double *>(scratch + lay.hii);
  CandRun *runs = reinterpret_cast<CandRun *>(scratch + lay.runs);
  u32 *cursor = reinterpret_cast<u32 *>(scratch + lay.cursor);
  u32 *cand = reinterpret_cast<u32 *>(scratch + lay.cand);
  const size_t smem = sizeof(OrbLists) + 16 * (size_t)table_offsets(g).total + sizeof(u32) * kCandPerWarp * kElocWarps;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(eloc_filter_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return check_launch("eloc_filter_kernel smem opt-in");
  const int w = CPLX ? 2 : 1;
  for (long long b0 = 0; b0 < n; b0 += lay.batch) {
    const long long nb = n - b0 < lay.batch ? n - b0 : lay.batch;
    const int splits = lay.splits;
    if (cudaMemsetAsync(cursor, 0, 4, st) != cudaSuccess) return check_launch("eloc cursor memset");
    eloc_filter_kernel<L><<<(unsigned)(nb * splits), kElocThreads, smem, st>>>(bra + b0 * L, nb, iv, runs, cand, cursor,
                                                                                (u32)lay.cand_cap, splits, g);
    count_launch();
    if (int rc = check_launch("eloc_filter_kernel")) return rc;
    const unsigned eb = (unsigned)((nb + kEvalThreads / 32 - 1) / (kEvalThreads / 32));
    eloc_eval_kernel<L, CPLX><<<eb, kEvalThreads, 0, st>>>(bra + b0 * L, nb, h1e, h2e, key, psi, N, iv, runs, cand, hii + b0,
                                                            eloc + b0 * w, psi0 + b0 * w, splits, g);
    count_launch();
    if (int rc = check_launch("eloc_eval_kernel")) return rc;
  }
  return 0;
}

int launch_eloc(const u64 *bra, long long n, const double *h1e, const double *h2e, const u64 *key, const double *psi, int cplx,
                long long N, const void *hash_ws, void *scratch, long long scratch_bytes, double *eloc, double *psi0,
                const ExcGeom &g, cudaStream_t st) {
  if (n == 0) return 0;
  const ElocScratch lay = eloc_scratch_layout(n, g.nsd);
  if (scratch_bytes < lay.total) {
    set_error("eloc scratch too small: %lld < %lld bytes", scratch_bytes, lay.total);
    return 4;
  }
  char *sc = static_cast<char *>(scratch);
  if (int rc = launch_diag_f64(bra, h1e, h2e, reinterpret_cast<double *>(sc + lay.hii), n, 1, g.L, g.sorb, g.nele, st)) return rc;
  const IndexView iv = index_view(hash_ws, N);
#define PYNQS_ELOC_CASE(LL)                                                                                                  \
  case LL:                                                                                                                   \
    return cplx ? launch_eloc_LC<LL, true>(bra, n, h1e, h2e, key, psi, N, iv, sc, lay, eloc, psi0, g, st)                    \
                : launch_eloc_LC<LL, false>(bra, n, h1e, h2e, key, psi, N, iv, sc, lay, eloc, psi0, g, st);
  switch (g.L) {
    PYNQS_ELOC_CASE(1)
    PYNQS_ELOC_CASE(2)
    PYNQS_ELOC_CASE(3)
  }
#undef PYNQS_ELOC_CASE
  set_error("unsupported ONV length L=%d", g.L);
  return 1;
}

}  // namespace pynqs
