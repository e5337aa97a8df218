// tables.cuh -- per-sample excitation tables in shared memory.
//
// Every connected determinant of a sample is  bra ^ m1 ^ m2  with m1, m2 drawn from six small
// per-sample tables (SURVEY.md App. A gives the index rules; cpp_src/cpu/excitation.cpp:18-110):
//     single alpha   SA[a*noA + i]          single beta   SB[b*noB + j]
//     hole pairs     HPA[ij], HPB[ij]        particle pairs PPA[ab], PPB[ab]
//     row r < d0: SA[r]            d0 <= r < d1: SB[r-d0]
//     d1 <= r < d2: HPA[r % noAA] ^ PPA[(r-d1)/noAA]        (r % noAA with the GLOBAL r: quirk Q1)
//     d2 <= r < d3: HPB[r % noBB] ^ PPB[(r-d2)/noBB]
//     r >= d3     : SA[q % sA] ^ SB[q / sA],  q = r - d3
// The tables hold C(na,2) entries per spin in total (noA*nvA + C(noA,2) + C(nvA,2) = C(na,2)),
// are built once per CTA by all threads, and replace the reference's per-row
// unpack_SinglesDoubles (15 div/mod + sqrt) by one exact multiply-shift division and two
// shared-memory loads.  Each entry also carries what the row needs next: the offset of its
// integral in the prepared tables and the sign bits (HitInfo below).
#pragma once
#include "common.cuh"

namespace pynqs {

struct TableOffsets {
  int sa, sb, hpa, ppa, hpb, ppb, total;
};

__host__ __device__ inline TableOffsets table_offsets(const ExcGeom &g) {
  TableOffsets t;
  t.sa = 0;
  t.sb = t.sa + g.sA;
  t.hpa = t.sb + g.noB * g.nvB;
  t.ppa = t.hpa + g.noAA;
  t.hpb = t.ppa + g.nvAA;
  t.ppb = t.hpb + g.noBB;
  t.total = t.ppb + g.nvBB;
  return t;
}

// Visit every table entry once: f(slot, kind, o0_entry, o1_entry) with kind 0 = SA, 1 = SB,
// 2 = HPA, 3 = PPA, 4 = HPB, 5 = PPB and the raw list entries (orbital | parity << 8).
// SA / SB: o0 = hole, o1 = particle; pair tables: o0 > o1.
template <typename F>
__device__ __forceinline__ void for_each_table_entry(const ExcGeom &g, const OrbLists &ls, const TableOffsets &to, F f) {
  for (int t = threadIdx.x; t < to.total; t += blockDim.x) {
    int kind, e0, e1;
    if (t < to.sb) {
      const u32 a = fdiv((u32)t, g.by_noA), i = (u32)t - a * g.noA;
      kind = 0; e0 = ls.a[i]; e1 = ls.a[g.noA + a];
    } else if (t < to.hpa) {
      const u32 u = (u32)(t - to.sb);
      const u32 b = fdiv(u, g.by_noB), j = u - b * g.noB;
      kind = 1; e0 = ls.b[j]; e1 = ls.b[g.noB + b];
    } else if (t < to.ppa) {
      int i, j;
      tri_unpack(t - to.hpa, i, j);
      kind = 2; e0 = ls.a[i]; e1 = ls.a[j];
    } else if (t < to.hpb) {
      int a, b;
      tri_unpack(t - to.ppa, a, b);
      kind = 3; e0 = ls.a[g.noA + a]; e1 = ls.a[g.noA + b];
    } else if (t < to.ppb) {
      int i, j;
      tri_unpack(t - to.hpb, i, j);
      kind = 4; e0 = ls.b[i]; e1 = ls.b[j];
    } else {
      int a, b;
      tri_unpack(t - to.ppb, a, b);
      kind = 5; e0 = ls.b[g.noB + a]; e1 = ls.b[g.noB + b];
    }
    f(t, kind, (u32)e0, (u32)e1);
  }
}

// ---- integral info of a table entry ------------------------------------------------------------
// off : offset contribution of the entry into the prepared tables (prepare.cu):
//         SA: [pA][hA] inner index of T_ab          SB: [hB][pB] outer index of T_ab
//         HP: column of T_aa / T_bb (hole pair)     PP: row (particle pair) * npair
//       bit 31 carries the entry's sign bit, so off1 + off2 yields offset and sign product in one
//       add (offsets stay below 2^28, no carry reaches bit 31).  Sign bit of a single = its own
//       sign par(h) ^ par(p) ^ [h < p]; of a pair = par(o0) ^ par(o1).
// cmp : orbitals packed for the "how many (hole, particle) pairs have hole < particle" term:
//         SA: hA | pA << 16      SB: pB | hB << 16      HP: hh | hl << 16      PP: ph | pl << 16
//       ((a | 0x80008000) - b has bit 15 = [a.lo >= b.lo] and bit 31 = [a.hi >= b.hi]).
struct HitInfo {
  u32 off, cmp;
};

__device__ __forceinline__ HitInfo make_hit_info(int kind, u32 e0, u32 e1, u32 na, u32 npair) {
  const u32 o0 = e0 & 0xffu, o1 = e1 & 0xffu, r0 = o0 >> 1, r1 = o1 >> 1;
  u32 sg = ((e0 ^ e1) >> 8) & 1u;
  if (kind < 2) sg ^= (u32)(o0 < o1);
  u32 off;
  if (kind == 0) off = r1 * na + r0;
  else if (kind == 1) off = (r0 * na + r1) * na * na;
  else if (kind == 2 || kind == 4) off = ((r0 * (r0 - 1)) >> 1) + r1;
  else off = (((r0 * (r0 - 1)) >> 1) + r1) * npair;
  HitInfo h;
  h.off = off | (sg << 31);
  h.cmp = kind == 1 ? (o1 | (o0 << 16)) : (o0 | (o1 << 16));
  return h;
}

// sign word (bit 31 = sign) of a double excitation from its two entries:
//   alpha-beta: s(SA) ^ s(SB) ^ [hA < pB] ^ [hB < pA] ^ 1 = s ^ [hA >= pB] ^ [pA >= hB]
//   same spin : s(HP) ^ s(PP) ^ #{(h, p): h < p} ^ 1      = s ^ 1 ^ XOR of the four [h >= p]
// (equal to parity(bra,h..) * parity(ket,p..) of cpp_src/cpu/excitation.cpp:153,161-163 because the
//  ket differs from the bra only in those four orbitals)
__device__ __forceinline__ u32 double_sign_word(bool alpha_beta, const HitInfo &i1, const HitInfo &i2) {
  const u32 sum = i1.off + i2.off;
  const u32 a = i1.cmp | 0x80008000u;
  if (alpha_beta) {
    const u32 d = a - i2.cmp;
    return sum ^ (d << 16) ^ d;
  }
  const u32 d = (a - (i2.cmp & 0xffffu) * 0x10001u) ^ (a - (i2.cmp >> 16) * 0x10001u);
  return ~(sum ^ (d << 16) ^ d);
}

// flip the sign of v when bit 31 of `bits` is set: exactly v * (+-1.0)
__device__ __forceinline__ double flip_sign(double v, u32 bits) {
  return __hiloint2double(__double2hiint(v) ^ (int)(bits & 0x80000000u), __double2loint(v));
}
__device__ __forceinline__ float flip_sign(float v, u32 bits) {
  return __int_as_float(__float_as_int(v) ^ (int)(bits & 0x80000000u));
}

}  // namespace pynqs
