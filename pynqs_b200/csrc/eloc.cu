// eloc.cu -- one-pass sample-space local energy.
//
// Fuses what the reference does in three native calls plus ~6 torch passes over [n, M]
// (vmc/energy/eloc.py:326-397: get_comb_hij_fused -> WavefunctionLUT.lookup -> scatter ->
// divide -> multiply -> sum): for every sample the CTA enumerates the connected determinants,
// probes the hash index of the sorted unique-sample table, and only for the hits evaluates
// <x|H|x'> and accumulates (psi(x')/psi(x)) * H.  Nothing of size [n, M] ever reaches HBM.
//
// Determinism: fixed row -> thread mapping, shuffle tree, ordered sum over warps and splits;
// no atomics.  Per element the arithmetic is the reference's (psi'/psi0 first, then * H);
// only the order of the final sum differs from torch's reduction (tolerance 1e-12 rel).
#include "lut.cuh"
#include "prepare.cuh"
#include "tables.cuh"

namespace pynqs {

constexpr int kElocThreads = 256;

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

// ---- table entries of the local-energy kernel ----------------------------------------------------
// aux: SA -> hash of the excited ALPHA string alpha(x) ^ m (tag / bucket inside a beta-grouped region)
//      SB -> descriptor of the region of the excited BETA string beta(x) ^ m (kNoRegion: no key of
//            the table has that beta string, so all noA*nvA alpha-beta doubles on top of it miss)
//      pair tables -> unused
template <int L>
struct __align__(16) LeEntry {
  u64 aux;
  u32 orbs, pad;  // o0 | o1 << 8
};
template <>
struct __align__(16) LeEntry<1> {
  u64 mask, aux;
};

template <int L>
__device__ __forceinline__ void le_set(LeEntry<L> &e, u32 o0, u32 o1) {
  e.orbs = o0 | (o1 << 8);
  e.pad = 0;
}
template <>
__device__ __forceinline__ void le_set<1>(LeEntry<1> &e, u32 o0, u32 o1) {
  e.mask = (1ull << o0) | (1ull << o1);
}
template <int L>
__device__ __forceinline__ Onv<L> le_apply(const Onv<L> &x, const LeEntry<L> &e) {
  Onv<L> y = x;
  flip_bit<L>(y, (int)(e.orbs & 0xffu));
  flip_bit<L>(y, (int)((e.orbs >> 8) & 0xffu));
  return y;
}
template <>
__device__ __forceinline__ Onv<1> le_apply<1>(const Onv<1> &x, const LeEntry<1> &e) {
  Onv<1> y;
  y.w[0] = x.w[0] ^ e.mask;
  return y;
}
// entry `index` of the table whose shared-window address is `saddr` (one LDS.128; ptxas narrows it
// when only the aux half is used)
template <int L>
__device__ __forceinline__ LeEntry<L> le_load(u32 saddr, int index) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr + 16u * (u32)index));
  LeEntry<L> e;
  *reinterpret_cast<uint4 *>(&e) = v;
  return e;
}

#ifndef PYNQS_ELOC_MIN_BLOCKS
#define PYNQS_ELOC_MIN_BLOCKS 5
#endif
constexpr int kElocMinBlocks = PYNQS_ELOC_MIN_BLOCKS;  // CTAs per SM the register allocation must allow
constexpr int kElocRows = 4;       // independent probes in flight per lane
constexpr int kHitQueueCap = 160;  // per-warp queue of found determinants (row, table index)

struct HitRec {
  u32 r, id;
};

// Per-CTA state in shared memory.  The slow paths (tag verification, hit evaluation) are
// out-of-line functions that read the geometry from HERE: handing them the kernel-parameter copy
// by reference would make the compiler spill it to local memory and reload it in the hot loop.
struct ElocShared {
  Cplx psi0;
  u64 desc_bx, desc_ax;  // regions of the sample's own beta / alpha string
  Cplx warp_sum[kElocThreads / 32];
  ExcGeom g;
  TableOffsets to;
  // operands of the slow paths (kept out of the hot loop's registers)
  const void *tab;
  const HitInfo *hit;
  const OrbLists *lists;
  const u64 *key;
  const double *h1e, *h2e, *psi;
  PrepView<double> prep;
};

// probe descriptor kept per SB entry / per sample: lo = first bucket of the region in the pool,
// hi = 32 - log2(buckets)  (so bucket = hash_hi >> hi and mask = 0xffffffff >> hi), hi = ~0: no region
__device__ __forceinline__ u64 pack_probe_desc(u64 region_desc) {
  if (region_desc == kNoRegion) return kNoRegion;
  return (region_desc & 0xffffffffull) | ((u64)(32u - (u32)(region_desc >> 32)) << 32);
}
__device__ __forceinline__ u32 shr_clamp(u32 v, u32 s) {  // PTX shr: shift amounts >= 32 give 0
  u32 r;
  asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
  return r;
}

template <int L>
__device__ __forceinline__ Onv<L> row_ket(const Onv<L> &x, const LeEntry<L> *tab, const ExcGeom &g, const TableOffsets &to, int r) {
  int t1, t2;
  const int cls = row_entries(g, to, r, t1, t2);
  Onv<L> y = le_apply<L>(x, tab[t1]);
  if (cls >= 2) y = le_apply<L>(y, tab[t2]);
  return y;
}

// region (probe-descriptor form) and string hash that row r has to be looked up with
template <int L>
__device__ __forceinline__ void row_probe_key(const Onv<L> &x, const ElocShared *sh, int r, u64 &desc, u64 &h2, Onv<L> &y) {
  const LeEntry<L> *tab = reinterpret_cast<const LeEntry<L> *>(sh->tab);
  int t1, t2;
  const int cls = row_entries(sh->g, sh->to, r, t1, t2);
  y = le_apply<L>(x, tab[t1]);
  if (cls >= 2) y = le_apply<L>(y, tab[t2]);
  if (cls == 4) {
    desc = tab[t2].aux;
    h2 = tab[t1].aux;
  } else if (cls == 0 || cls == 2) {
    desc = sh->desc_bx;
    h2 = cls == 0 ? tab[t1].aux : hash_alpha<L>(y);
  } else {
    desc = sh->desc_ax;
    h2 = hash_beta<L>(y);
  }
}

// Out of line (keeps the probe loop small for the instruction cache and its registers few): the
// first bucket `tj` of row r's probe holds a matching tag or has overflowed.  Re-derive the probe
// from r, compare candidates with the key table, follow overflowed buckets.  Table row or -1.
template <int L>
__device__ __noinline__ int eloc_resolve(Onv<L> x, const ElocShared *sh, const HashBucket *__restrict__ pool, int r, uint4 tj) {
  u64 desc, h2;
  Onv<L> y;
  row_probe_key<L>(x, sh, r, desc, h2, y);
  const u32 roff = (u32)desc, sh32 = (u32)(desc >> 32), msk = shr_clamp(~0u, sh32), tag = hash_tag(h2);
  u32 b = shr_clamp((u32)(h2 >> 32), sh32);
  const u64 *__restrict__ key = sh->key;
#pragma unroll 1
  for (u32 probe = 0;; ++probe) {
    if (tags_match(tj, tag)) {
      const u32 tg[4] = {tj.x, tj.y, tj.z, tj.w | 1u};
#pragma unroll 1
      for (int sl = 0; sl < 4; ++sl) {
        if (tg[sl] != tag) continue;
        const u32 cand = __ldg(&pool[roff + b].idx[sl]);
        if (eq_onv<L>(load_onv<L>(key + (long long)cand * L), y)) return (int)cand;
      }
    }
    if (!bucket_overflowed(tj) || probe >= msk) return -1;
    b = (b + 1) & msk;
    tj = __ldg(reinterpret_cast<const uint4 *>(pool[roff + b].tag));
  }
}

// <x|H|x'> of excitation r through the per-sample tables and the prepared integrals -- the same
// numbers, order of additions and sign as the fused operator (enumerate.cu); falls back to the
// packed arrays when no prepared workspace was given.
template <int L>
__device__ __forceinline__ double row_element(const Onv<L> &x, const HitInfo *hit, const OrbLists *lists, const ExcGeom &g,
                                              const TableOffsets &to, int r, const double *__restrict__ h1e,
                                              const double *__restrict__ h2e, const PrepView<double> &prep) {
  if (prep.ab == nullptr) return exc_element<L, double>(x, decode_exc(g, *lists, r), h1e, h2e, g.sorb);
  int t1, t2;
  const int cls = row_entries(g, to, r, t1, t2);
  const HitInfo i1 = hit[t1];
  if (cls >= 2) {
    const HitInfo i2 = hit[t2];
    const double *tbl = cls == 4 ? prep.ab : (cls == 2 ? prep.aa : prep.bb);
    return flip_sign(1.0 * __ldg(tbl + ((i1.off + i2.off) & 0x7fffffffu)), double_sign_word(cls == 4, i1, i2));
  }
  const u32 h = cls == 0 ? (i1.cmp & 0xffu) : (i1.cmp >> 16), p = cls == 0 ? (i1.cmp >> 16) : (i1.cmp & 0xffu);
  const u32 na = (u32)prep.na;
  const size_t kstride = (size_t)2 * na * na;
  const int n_occ = lists->n_occ;
  double v = 0.0;
  v += __ldg(h1e + (size_t)p * g.sorb + h);
  const double *line = prep.s + ((size_t)(h & 1u) * na + (p >> 1)) * na + (h >> 1);
  for (int q = 0; q < n_occ; ++q) v += __ldg(line + kstride * lists->occ_order[q]);
  return flip_sign(v, i1.off);
}

// drain a warp's queue with all lanes busy: lane e evaluates hit e.  Returns the lane's sum of
// (psi'/psi0) * <x|H|x'> over its hits.
template <int L, bool CPLX>
__device__ __noinline__ Cplx eloc_flush(Onv<L> x, const ElocShared *sh, const HitRec *queue, u32 count) {
  Cplx acc = {0.0, 0.0};
  const Cplx p0 = sh->psi0;
  const double *__restrict__ psi = sh->psi;
  __syncwarp();
  for (u32 e = threadIdx.x & 31; e < count; e += 32) {
    const HitRec h = queue[e];
    const double hval = row_element<L>(x, sh->hit, sh->lists, sh->g, sh->to, (int)h.r, sh->h1e, sh->h2e, sh->prep);
    accumulate<CPLX>(acc, load_psi<CPLX>(psi, (long long)h.id), p0, hval);
  }
  __syncwarp();
  return acc;
}

// One chunk of 32 * ROWS consecutive rows of one excitation class (CLS 0/1 single a/b, 2/3 double
// aa/bb, 4 double ab; excitation index r = row), lane l owning rows c + l + 32 j.  Phase 1 issues the
// first bucket load of ROWS probes per lane; phase 2 checks the four tags of each bucket -- a probe
// ends there unless a tag matches or the bucket has overflowed (rare, out of line).  Found
// determinants are compacted (ballot + popcount prefix) into the warp's queue and evaluated later
// by full warps.  FULL: every row of the chunk is inside [lo, hi).
template <int L, int CLS, int ROWS, bool FULL>
__device__ __forceinline__ void eloc_chunk(const Onv<L> &x, u32 tab_s, const TableOffsets &to, const ExcGeom &g, u64 own,
                                           const ElocShared *sh, const HashBucket *__restrict__ pool, HitRec *queue, u32 &qn,
                                           int c, int hi) {
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
        h2 = le_load<L>(tab_s, to.sa + r).aux;
      } else if (CLS == 1) {
        desc = own;
        h2 = hash_beta<L>(le_apply<L>(x, le_load<L>(tab_s, to.sb + (r - g.d0))));
      } else if (CLS == 2) {
        const LeEntry<L> e1 = le_load<L>(tab_s, to.hpa + (int)((u32)r - fdiv((u32)r, g.by_noAA) * g.noAA));  // global r (quirk Q1)
        const LeEntry<L> e2 = le_load<L>(tab_s, to.ppa + (int)fdiv((u32)(r - g.d1), g.by_noAA));
        desc = own;
        h2 = hash_alpha<L>(le_apply<L>(le_apply<L>(x, e1), e2));
      } else if (CLS == 3) {
        const LeEntry<L> e1 = le_load<L>(tab_s, to.hpb + (int)((u32)r - fdiv((u32)r, g.by_noBB) * g.noBB));
        const LeEntry<L> e2 = le_load<L>(tab_s, to.ppb + (int)fdiv((u32)(r - g.d2), g.by_noBB));
        desc = own;
        h2 = hash_beta<L>(le_apply<L>(le_apply<L>(x, e1), e2));
      } else {
        const u32 q = (u32)(r - g.d3);
        const u32 jb = fdiv(q, g.by_sA);
        h2 = le_load<L>(tab_s, to.sa + (int)(q - jb * g.sA)).aux;  // hash of the excited alpha string
        desc = le_load<L>(tab_s, to.sb + (int)jb).aux;             // probe descriptor of the excited beta string's region
      }
    }
    const u32 sh32 = (u32)(desc >> 32);
    tag[j] = 0;
    if (sh32 != ~0u) {
      tag[j] = hash_tag(h2);
      t[j] = __ldg(reinterpret_cast<const uint4 *>(pool[(u32)desc + shr_clamp((u32)(h2 >> 32), sh32)].tag));
    }
  }
  int id[ROWS];
  bool any = false;
#pragma unroll
  for (int j = 0; j < ROWS; ++j) {
    id[j] = -1;
    if (tag[j] && (tags_match(t[j], tag[j]) || bucket_overflowed(t[j]))) {
      id[j] = eloc_resolve<L>(x, sh, pool, c + lane + 32 * j, t[j]);
      any |= id[j] >= 0;
    }
  }
  if (__any_sync(0xffffffffu, any)) {
    const u32 lt = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      const u32 m = __ballot_sync(0xffffffffu, id[j] >= 0);
      if (id[j] >= 0) {
        HitRec h;
        h.r = (u32)(c + lane + 32 * j);
        h.id = (u32)id[j];
        queue[qn + __popc(m & lt)] = h;
      }
      qn += __popc(m);
    }
  }
}

template <int L, bool CPLX, int CLS, int ROWS>
__device__ __forceinline__ void eloc_class(const Onv<L> &x, u32 tab_s, const TableOffsets &to, const ExcGeom &g,
                                           const ElocShared *sh, const HashBucket *__restrict__ pool, HitRec *queue, u32 &qn,
                                           Cplx &acc, int lo, int hi) {
  constexpr bool kBetaGrouped = (CLS == 0 || CLS == 2 || CLS == 4);
  const u64 own = kBetaGrouped ? sh->desc_bx : sh->desc_ax;  // already in probe-descriptor form
  if (CLS != 4 && own == kNoRegion) return;  // the sample's own string is not in the table: nothing of this class is
  constexpr int kChunk = 32 * ROWS;
  for (int c = lo + (int)(threadIdx.x >> 5) * kChunk; c < hi; c += (kElocThreads / 32) * kChunk) {
    if (c + kChunk <= hi) eloc_chunk<L, CLS, ROWS, true>(x, tab_s, to, g, own, sh, pool, queue, qn, c, hi);
    else eloc_chunk<L, CLS, ROWS, false>(x, tab_s, to, g, own, sh, pool, queue, qn, c, hi);
    if (qn > (u32)(kHitQueueCap - kChunk)) {
      const Cplx part = eloc_flush<L, CPLX>(x, sh, queue, qn);
      acc.re += part.re;
      acc.im += part.im;
      qn = 0;
    }
  }
}

// tables with duplicate keys: the reference's classic search for every row (rare, out of line)
template <int L, bool CPLX>
__device__ __noinline__ Cplx eloc_rows_classic(Onv<L> x, const ElocShared *sh, long long N, int r_begin, int r_end) {
  Cplx acc = {0.0, 0.0};
  const LeEntry<L> *tab = reinterpret_cast<const LeEntry<L> *>(sh->tab);
  for (int r = r_begin + threadIdx.x; r < r_end; r += kElocThreads) {
    const long long id = classic_search<L>(sh->key, N, row_ket<L>(x, tab, sh->g, sh->to, r));
    if (id >= 0)
      accumulate<CPLX>(acc, load_psi<CPLX>(sh->psi, id), sh->psi0,
                       row_element<L>(x, sh->hit, sh->lists, sh->g, sh->to, r, sh->h1e, sh->h2e, sh->prep));
  }
  return acc;
}

// the sample's own lookup: psi0 and the regions of its two strings (one thread per CTA, out of line)
template <int L, bool CPLX>
__device__ __noinline__ void eloc_sample_setup(Onv<L> x, ElocShared *sh, IndexView iv, long long N, bool dup) {
  const u64 dbx = dir_find(iv.dir[0], iv.log2_dir, hash_beta<L>(x));
  sh->desc_bx = pack_probe_desc(dbx);
  sh->desc_ax = pack_probe_desc(dir_find(iv.dir[1], iv.log2_dir, hash_alpha<L>(x)));
  long long id = -1;
  if (dup) id = classic_search<L>(sh->key, N, x);
  else if (dbx != kNoRegion) id = region_probe<L>(sh->key, iv.pool, dbx, hash_alpha<L>(x), [&]() { return x; });
  Cplx p0 = {0.0, 0.0};
  if (id >= 0) p0 = load_psi<CPLX>(sh->psi, id);
  sh->psi0 = p0;
}

template <int L, bool CPLX>
__global__ void __launch_bounds__(kElocThreads, kElocMinBlocks)
eloc_kernel(const u64 *__restrict__ bra, long long n, const double *__restrict__ h1e, const double *__restrict__ h2e,
            PrepView<double> prep, const u64 *__restrict__ key, const double *__restrict__ psi, long long N, IndexView iv,
            const double *__restrict__ hii, double *__restrict__ partial, double *__restrict__ psi0_out, int splits,
            ExcGeom g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const TableOffsets to = table_offsets(g);
  OrbLists &lists = *reinterpret_cast<OrbLists *>(smem_raw);
  LeEntry<L> *tab = reinterpret_cast<LeEntry<L> *>(smem_raw + sizeof(OrbLists));
  HitInfo *hit = reinterpret_cast<HitInfo *>(tab + to.total);
  HitRec *queues = reinterpret_cast<HitRec *>(hit + to.total);
  __shared__ ElocShared sh;

  const long long s = blockIdx.x / splits;
  const int split = blockIdx.x - (int)(s * splits);
  if (s >= n) return;
  const Onv<L> x = load_onv<L>(bra + s * L);
  const bool dup = iv.hdr->has_dup != 0;  // table with duplicate keys: classic search (reference probe sequence)
  if (threadIdx.x < 32) build_lists<L>(x, g.sorb, g.noA, g.noB, lists, threadIdx.x);
  if (threadIdx.x == 32) {
    sh.g = g;
    sh.to = to;
    sh.tab = tab;
    sh.hit = hit;
    sh.lists = &lists;
    sh.key = key;
    sh.h1e = h1e;
    sh.h2e = h2e;
    sh.psi = psi;
    sh.prep = prep;
    eloc_sample_setup<L, CPLX>(x, &sh, iv, N, dup);
    if (split == 0) {
      psi0_out[CPLX ? 2 * s : s] = sh.psi0.re;
      if (CPLX) psi0_out[2 * s + 1] = sh.psi0.im;
    }
  }
  __syncthreads();
  const u32 na = (u32)(g.sorb / 2), npair = na * (na - 1) / 2;
  for_each_table_entry(g, lists, to, [&](int t, int kind, u32 e0, u32 e1) {
    LeEntry<L> e;
    le_set<L>(e, e0 & 0xffu, e1 & 0xffu);
    e.aux = 0;
    if (!dup) {
      if (kind == 0) e.aux = hash_alpha<L>(le_apply<L>(x, e));
      else if (kind == 1) e.aux = pack_probe_desc(dir_find(iv.dir[0], iv.log2_dir, hash_beta<L>(le_apply<L>(x, e))));
    }
    tab[t] = e;
    hit[t] = make_hit_info(kind, e0, e1, na, npair);
  });
  __syncthreads();
  const Cplx p0 = sh.psi0;

  Cplx acc = {0.0, 0.0};
  if (split == 0 && threadIdx.x == 0) accumulate<CPLX>(acc, p0, p0, hii[s]);  // row 0: (psi0/psi0) * H_xx

  const int chunk = (g.nsd + splits - 1) / splits;
  const int r_begin = split * chunk;
  const int r_end = min(g.nsd, r_begin + chunk);
  if (dup) {
    const Cplx part = eloc_rows_classic<L, CPLX>(x, &sh, N, r_begin, r_end);
    acc.re += part.re;
    acc.im += part.im;
  } else {
    HitRec *queue = queues + (threadIdx.x >> 5) * kHitQueueCap;
    u32 qn = 0;
    const u32 tab_s = (u32)__cvta_generic_to_shared(tab);
    auto lo = [&](int v) { return v > r_begin ? v : r_begin; };
    auto hi = [&](int v) { return v < r_end ? v : r_end; };
    eloc_class<L, CPLX, 0, 1>(x, tab_s, to, g, &sh, iv.pool, queue, qn, acc, lo(0), hi(g.d0));
    eloc_class<L, CPLX, 1, 1>(x, tab_s, to, g, &sh, iv.pool, queue, qn, acc, lo(g.d0), hi(g.d1));
    eloc_class<L, CPLX, 2, kElocRows>(x, tab_s, to, g, &sh, iv.pool, queue, qn, acc, lo(g.d1), hi(g.d2));
    eloc_class<L, CPLX, 3, kElocRows>(x, tab_s, to, g, &sh, iv.pool, queue, qn, acc, lo(g.d2), hi(g.d3));
    eloc_class<L, CPLX, 4, kElocRows>(x, tab_s, to, g, &sh, iv.pool, queue, qn, acc, lo(g.d3), hi(g.nsd));
    if (qn) {
      const Cplx part = eloc_flush<L, CPLX>(x, &sh, queue, qn);
      acc.re += part.re;
      acc.im += part.im;
    }
  }

  // deterministic block reduction
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc.re += __shfl_xor_sync(0xffffffffu, acc.re, o);
    if (CPLX) acc.im += __shfl_xor_sync(0xffffffffu, acc.im, o);
  }
  if ((threadIdx.x & 31) == 0) sh.warp_sum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    Cplx t = sh.warp_sum[0];
#pragma unroll
    for (int w = 1; w < kElocThreads / 32; ++w) {
      t.re += sh.warp_sum[w].re;
      t.im += sh.warp_sum[w].im;
    }
    const long long o = s * splits + split;
    if (CPLX) {
      partial[2 * o] = t.re;
      partial[2 * o + 1] = t.im;
    } else {
      partial[o] = t.re;
    }
  }
}

template <bool CPLX>
__global__ void eloc_finish_kernel(const double *__restrict__ partial, double *__restrict__ eloc, long long n, int splits) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  double re = 0.0, im = 0.0;
  for (int k = 0; k < splits; ++k) {
    if (CPLX) {
      re += partial[2 * (s * splits + k)];
      im += partial[2 * (s * splits + k) + 1];
    } else {
      re += partial[s * splits + k];
    }
  }
  if (CPLX) {
    eloc[2 * s] = re;
    eloc[2 * s + 1] = im;
  } else {
    eloc[s] = re;
  }
}

int eloc_splits(long long n, int nsd) {
  if (n <= 0) return 1;
  long long want = (148LL * 8 + n - 1) / n;
  long long cap = (nsd + 2047) / 2048;
  if (cap < 1) cap = 1;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

long long eloc_scratch_bytes(long long n, int nsd, int cplx) {
  const int splits = eloc_splits(n, nsd);
  const long long w = cplx ? 2 : 1;
  // hii[n] | partial[n * splits * w]
  return 8 * (n + n * splits * w) + 256;
}

int launch_diag_f64(const u64 *bra, const double *h1e, const double *h2e, double *out, long long n, long long stride, int L,
                    int sorb, int nele, cudaStream_t st);

template <int L>
static int launch_eloc_L(const u64 *bra, long long n, const double *h1e, const double *h2e, const void *prep_ws, const u64 *key,
                         const double *psi,
                         int cplx, long long N, const IndexView &hdr, double *hii, double *partial, double *eloc,
                         double *psi0, int splits, const ExcGeom &g, cudaStream_t st) {
  const long long blocks = n * splits;
  if (blocks > 0x7fffffffLL) {
    set_error("eloc: n * splits = %lld exceeds the grid limit; split the batch", blocks);
    return 1;
  }
  double *dst = splits == 1 ? eloc : partial;
  const size_t entries = (size_t)table_offsets(g).total;
  const size_t smem = sizeof(OrbLists) + (sizeof(LeEntry<L>) + sizeof(HitInfo)) * entries + sizeof(HitRec) * kHitQueueCap * (kElocThreads / 32);
  PrepView<double> pv;
  if (prep_ws) pv = prep_view<double>(prep_ws, g.sorb);
  else pv.ab = pv.aa = pv.bb = pv.s = nullptr, pv.na = g.sorb / 2, pv.npair = 0;
  if (smem > 48 * 1024) {
    cudaError_t e = cplx ? cudaFuncSetAttribute(eloc_kernel<L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                         : cudaFuncSetAttribute(eloc_kernel<L, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_launch("eloc_kernel smem opt-in");
  }
  if (cplx)
    eloc_kernel<L, true><<<(unsigned)blocks, kElocThreads, smem, st>>>(bra, n, h1e, h2e, pv, key, psi, N, hdr, hii, dst, psi0, splits, g);
  else
    eloc_kernel<L, false><<<(unsigned)blocks, kElocThreads, smem, st>>>(bra, n, h1e, h2e, pv, key, psi, N, hdr, hii, dst, psi0, splits, g);
  count_launch();
  if (int rc = check_launch("eloc_kernel")) return rc;
  if (splits > 1) {
    const unsigned fb = (unsigned)((n + 255) / 256);
    if (cplx) eloc_finish_kernel<true><<<fb, 256, 0, st>>>(partial, eloc, n, splits);
    else eloc_finish_kernel<false><<<fb, 256, 0, st>>>(partial, eloc, n, splits);
    count_launch();
    if (int rc = check_launch("eloc_finish_kernel")) return rc;
  }
  return 0;
}

int launch_eloc(const u64 *bra, long long n, const double *h1e, const double *h2e, const void *prep_ws, const u64 *key,
                const double *psi, int cplx, long long N, const void *hash_ws, void *scratch, long long scratch_bytes, double *eloc, double *psi0,
                const ExcGeom &g, cudaStream_t st) {
  if (n == 0) return 0;
  const long long need = eloc_scratch_bytes(n, g.nsd, cplx);
  if (scratch_bytes < need) {
    set_error("eloc scratch too small: %lld < %lld bytes", scratch_bytes, need);
    return 4;
  }
  const int splits = eloc_splits(n, g.nsd);
  double *hii = reinterpret_cast<double *>(scratch);
  double *partial = hii + n;
  if (int rc = launch_diag_f64(bra, h1e, h2e, hii, n, 1, g.L, g.sorb, g.nele, st)) return rc;
  const IndexView hdr = index_view(hash_ws, N);
  switch (g.L) {
    case 1: return launch_eloc_L<1>(bra, n, h1e, h2e, prep_ws, key, psi, cplx, N, hdr, hii, partial, eloc, psi0, splits, g, st);
    case 2: return launch_eloc_L<2>(bra, n, h1e, h2e, prep_ws, key, psi, cplx, N, hdr, hii, partial, eloc, psi0, splits, g, st);
    case 3: return launch_eloc_L<3>(bra, n, h1e, h2e, prep_ws, key, psi, cplx, N, hdr, hii, partial, eloc, psi0, splits, g, st);
  }
  set_error("unsupported ONV length L=%d", g.L);
  return 1;
}

}  // namespace pynqs
