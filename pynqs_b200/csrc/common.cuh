// common.cuh -- shared device helpers for the sm_100a local-energy kernels.
//
// Semantics restated from the reference (paths under /root/reference/):
//   merged occupied/virtual lists   cpp_src/cpu/onstate.cpp:147-193
//   flat index -> excitation        cpp_src/cpu/excitation.cpp:18-110, excitation.h:6-11
//   packed integrals                cpp_src/cpu/hamiltonian.cpp:7-31
//   Slater-Condon elements          cpp_src/cpu/hamiltonian.cpp:33-102, excitation.cpp:124-169
// The arithmetic ORDER of every floating-point sum is the reference's, so H_ij is bit-identical.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pynqs {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int kMaxL = 3;
constexpr int kMaxHalf = 96;  // orbitals per spin at 192 spin orbitals

// ---- exact unsigned division by a launch-uniform divisor ------------------------------------
// q = umulhi(n, mul) >> shift is exact for every n < 2^31 (mul = ceil(2^(32+shift)/d),
// 2^shift < d <= 2^(shift+1)); d == 1 is flagged with mul == 0.
struct FastDiv {
  u32 mul, shift, d, pad;
};

inline FastDiv make_fastdiv(u32 d) {
  FastDiv f;
  f.d = d;
  f.pad = 0;
  if (d <= 1) {
    f.mul = 0;
    f.shift = 0;
    return f;
  }
  u32 s = 0;
  while ((1ull << (s + 1)) < d) ++s;  // 2^s < d <= 2^(s+1)
  unsigned __int128 num = (unsigned __int128)1 << (32 + s);
  f.mul = (u32)((num + d - 1) / d);
  f.shift = s;
  return f;
}

__device__ __forceinline__ u32 fdiv(u32 n, const FastDiv &f) {
  return f.mul ? (__umulhi(n, f.mul) >> f.shift) : n;
}

// ---- per-launch excitation geometry (identical for every sample) ----------------------------
struct ExcGeom {
  int sorb, L, nele;
  int noA, noB, nvA, nvB;
  int noAA, noBB, nvAA, nvBB;
  int d0, d1, d2, d3, nsd;  // block ends: S_a, S_b, D_aa, D_bb, then D_ab up to nsd
  int sA;                   // noA * nvA
  FastDiv by_noA, by_noB, by_noAA, by_noBB, by_sA;
};

inline ExcGeom make_geom(int sorb, int nele, int noA, int noB) {
  ExcGeom g;
  g.sorb = sorb;
  g.L = (sorb - 1) / 64 + 1;
  g.nele = nele;
  int k = sorb / 2;
  g.noA = noA;
  g.noB = noB;
  g.nvA = k - noA;
  g.nvB = k - noB;
  g.noAA = noA * (noA - 1) / 2;
  g.noBB = noB * (noB - 1) / 2;
  g.nvAA = g.nvA * (g.nvA - 1) / 2;
  g.nvBB = g.nvB * (g.nvB - 1) / 2;
  g.sA = noA * g.nvA;
  g.d0 = g.sA;
  g.d1 = g.d0 + noB * g.nvB;
  g.d2 = g.d1 + g.noAA * g.nvAA;
  g.d3 = g.d2 + g.noBB * g.nvBB;
  g.nsd = g.d3 + g.sA * noB * g.nvB;
  g.by_noA = make_fastdiv((u32)(noA > 0 ? noA : 1));
  g.by_noB = make_fastdiv((u32)(noB > 0 ? noB : 1));
  g.by_noAA = make_fastdiv((u32)(g.noAA > 0 ? g.noAA : 1));
  g.by_noBB = make_fastdiv((u32)(g.noBB > 0 ? g.noBB : 1));
  g.by_sA = make_fastdiv((u32)(g.sA > 0 ? g.sA : 1));
  return g;
}

// ---- bit helpers ------------------------------------------------------------------------------
template <int L>
struct Onv {
  u64 w[L];
};

template <int L>
__device__ __forceinline__ Onv<L> load_onv(const u64 *p) {
  Onv<L> x;
#pragma unroll
  for (int i = 0; i < L; ++i) x.w[i] = p[i];
  return x;
}

template <int L>
__device__ __forceinline__ bool test_bit(const Onv<L> &x, int k) {
  u64 w = x.w[0];
  if (L > 1) {
#pragma unroll
    for (int i = 1; i < L; ++i)
      if ((k >> 6) == i) w = x.w[i];
  }
  return (w >> (k & 63)) & 1ull;
}

template <int L>
__device__ __forceinline__ void flip_bit(Onv<L> &x, int k) {
  if (L == 1) {
    x.w[0] ^= 1ull << k;
  } else {
#pragma unroll
    for (int i = 0; i < L; ++i)
      if ((k >> 6) == i) x.w[i] ^= 1ull << (k & 63);
  }
}

// mask of word i restricted to bits of global index < k
__device__ __forceinline__ u64 below_in_word(int k, int i) {
  int r = k - 64 * i;
  return r <= 0 ? 0ull : (r >= 64 ? ~0ull : ((1ull << r) - 1ull));
}

// number of occupied orbitals strictly below k   (parity of this = cpp_src/cpu/onstate.cpp:22-32)
template <int L>
__device__ __forceinline__ int count_below(const Onv<L> &x, int k) {
  int c = 0;
#pragma unroll
  for (int i = 0; i < L; ++i) c += __popcll(x.w[i] & below_in_word(k, i));
  return c;
}

constexpr u64 kEven = 0x5555555555555555ull;  // alpha spin orbitals
constexpr u64 kOdd = 0xAAAAAAAAAAAAAAAAull;   // beta spin orbitals

// ---- per-sample orbital lists in shared memory -------------------------------------------------
// entry = orbital | (parity of the bra's occupation below that orbital) << 8.
// lstA[t]: t-th alpha orbital, occupied ones first (ascending) then virtual (ascending); lstB same
// for beta.  This is the reference's merged list de-interleaved: merged[2t] = lstA[t],
// merged[2t+1] = lstB[t]   (cpp_src/cpu/onstate.cpp:147-193).
struct OrbLists {
  unsigned short a[kMaxHalf];
  unsigned short b[kMaxHalf];
  // occupied orbitals in the order the reference sums a single excitation's two-electron terms:
  // words ascending, bits DEscending inside a word (cpp_src/cpu/hamiltonian.cpp:59-68)
  unsigned char occ_order[2 * kMaxHalf];
  int n_occ;
  int pad[3];
};

// One warp builds both lists: lane handles orbitals lane, lane+32, ...  The bra word itself is
// the occupancy ballot; ranks come from popcounts of masked words.
template <int L>
__device__ __forceinline__ void build_lists(const Onv<L> &x, int sorb, int noA, int noB, OrbLists &out, int lane, int stride = 32) {
  for (int k = lane; k < sorb; k += stride) {
    const bool occ = test_bit<L>(x, k);
    const u64 spin = (k & 1) ? kOdd : kEven;
    int occ_same = 0, occ_all = 0, below_same = 0;
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const u64 m = below_in_word(k, j);
      occ_same += __popcll(x.w[j] & spin & m);
      occ_all += __popcll(x.w[j] & m);
      below_same += __popcll(spin & m);
    }
    const int no = (k & 1) ? noB : noA;
    const int slot = occ ? occ_same : no + (below_same - occ_same);
    const unsigned short e = (unsigned short)(k | ((occ_all & 1) << 8));
    if (k & 1) out.b[slot] = e;
    else out.a[slot] = e;
    if (occ) {
      // position in the single-excitation summation order
      int pos = 0;
#pragma unroll
      for (int j = 0; j < L; ++j) {
        if (j < (k >> 6)) pos += __popcll(x.w[j]);
        else if (j == (k >> 6)) pos += __popcll(x.w[j] >> (k & 63)) - 1;  // occupied bits above k in its word
      }
      out.occ_order[pos] = (unsigned char)k;
    }
  }
  if (lane == 0) {
    int c = 0;
#pragma unroll
    for (int j = 0; j < L; ++j) c += __popcll(x.w[j]);
    out.n_occ = c;
  }
}

// triangular pair index t -> (hi, lo), hi > lo >= 0.  The reference computes
// hi = int(sqrt(2(t+1)) + 0.5) in FP64 (excitation.h:6-11); 2(t+1) lies in
// [hi^2-hi+2, hi^2+hi], strictly inside ((hi-.5)^2, (hi+.5)^2), so any correctly rounded
// single-precision sqrt (even the approximate MUFU one) followed by round-to-nearest yields the same
// integer for hi < 2^11.
__device__ __forceinline__ void tri_unpack(int t, int &hi, int &lo) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"((float)(2 * t + 2)));  // error ~1e-6 relative, margin ~1e-2
  hi = __float2int_rn(r);
  lo = t - ((hi * (hi - 1)) >> 1);
}

// One decoded excitation: orbitals (with parity bit 8) of holes h0,h1 and particles p0,p1.
struct Exc {
  int h0, p0, h1, p1;  // raw list entries; singles leave h1/p1 unused
  bool dbl;
};

// flat index r in [0, nsd) -> excitation   (cpp_src/cpu/excitation.cpp:18-110, quirk: same-spin
// doubles take the hole pair from r % noAA with the GLOBAL r).
__device__ __forceinline__ Exc decode_exc(const ExcGeom &g, const OrbLists &ls, int r) {
  Exc e;
  if (r < g.d0) {
    u32 q = fdiv((u32)r, g.by_noA);
    u32 i = (u32)r - q * g.noA;
    e.h0 = ls.a[i];
    e.p0 = ls.a[g.noA + q];
    e.h1 = e.p1 = 0;
    e.dbl = false;
  } else if (r < g.d1) {
    u32 t = (u32)(r - g.d0);
    u32 q = fdiv(t, g.by_noB);
    u32 i = t - q * g.noB;
    e.h0 = ls.b[i];
    e.p0 = ls.b[g.noB + q];
    e.h1 = e.p1 = 0;
    e.dbl = false;
  } else if (r < g.d2) {
    u32 ab = fdiv((u32)(r - g.d1), g.by_noAA);
    u32 ij = (u32)r - fdiv((u32)r, g.by_noAA) * g.noAA;
    int i, j, a, b;
    tri_unpack((int)ij, i, j);
    tri_unpack((int)ab, a, b);
    e.h0 = ls.a[i];
    e.h1 = ls.a[j];
    e.p0 = ls.a[g.noA + a];
    e.p1 = ls.a[g.noA + b];
    e.dbl = true;
  } else if (r < g.d3) {
    u32 ab = fdiv((u32)(r - g.d2), g.by_noBB);
    u32 ij = (u32)r - fdiv((u32)r, g.by_noBB) * g.noBB;
    int i, j, a, b;
    tri_unpack((int)ij, i, j);
    tri_unpack((int)ab, a, b);
    e.h0 = ls.b[i];
    e.h1 = ls.b[j];
    e.p0 = ls.b[g.noB + a];
    e.p1 = ls.b[g.noB + b];
    e.dbl = true;
  } else {
    u32 t = (u32)(r - g.d3);
    u32 jb = fdiv(t, g.by_sA);
    u32 ia = t - jb * g.sA;
    u32 a = fdiv(ia, g.by_noA), i = ia - a * g.noA;
    u32 b = fdiv(jb, g.by_noB), j = jb - b * g.noB;
    e.h0 = ls.a[i];
    e.p0 = ls.a[g.noA + a];
    e.h1 = ls.b[j];
    e.p1 = ls.b[g.noB + b];
    e.dbl = true;
  }
  return e;
}

// ket = bra with the excitation's orbitals flipped
template <int L>
__device__ __forceinline__ Onv<L> apply_exc(const Onv<L> &x, const Exc &e) {
  Onv<L> y = x;
  flip_bit<L>(y, e.h0 & 0xff);
  flip_bit<L>(y, e.p0 & 0xff);
  if (e.dbl) {
    flip_bit<L>(y, e.h1 & 0xff);
    flip_bit<L>(y, e.p1 & 0xff);
  }
  return y;
}

// ---- packed integrals ---------------------------------------------------------------------------
// <ij||kl> with sign bookkeeping of cpp_src/cpu/hamiltonian.cpp:13-31
template <typename T>
__device__ __forceinline__ T two_body(const T *__restrict__ h2e, u32 i, u32 j, u32 k, u32 l) {
  if (i == j || k == l) return (T)0.0;
  const u32 ij = i > j ? ((i * (i - 1)) >> 1) + j : ((j * (j - 1)) >> 1) + i;
  const u32 kl = k > l ? ((k * (k - 1)) >> 1) + l : ((l * (l - 1)) >> 1) + k;
  T s = (T)1.0;
  if (!(i > j)) s = -s;
  if (!(k > l)) s = -s;
  const u32 hi = ij >= kl ? ij : kl, lo = ij >= kl ? kl : ij;
  const size_t off = (((size_t)hi * (hi + 1)) >> 1) + lo;
  return s * __ldg(h2e + off);
}

// offset of <p0 p1||q0 q1>, p0 > p1, q0 > q1 (sign +)
__device__ __forceinline__ size_t pair_offset(u32 p0, u32 p1, u32 q0, u32 q1) {
  const u32 ij = ((p0 * (p0 - 1)) >> 1) + p1;
  const u32 kl = ((q0 * (q0 - 1)) >> 1) + q1;
  const u32 hi = ij >= kl ? ij : kl, lo = ij >= kl ? kl : ij;
  return (((size_t)hi * (hi + 1)) >> 1) + lo;
}

// Element of a decoded excitation.  Parity bits ride in bit 8 of the list entries:
//   single: sign = par(h) ^ par(p) ^ [h < p]
//   double: sign = par(h0)^par(h1)^par(p0)^par(p1) ^ #{(h,p): h < p} ^ 1
// (equals parity(bra,h..)*parity(ket,p..) of excitation.cpp:153,161-163 because the ket differs
//  from the bra only in those four orbitals).
template <int L, typename T>
__device__ __forceinline__ T exc_element(const Onv<L> &x, const Exc &e, const T *__restrict__ h1e,
                                         const T *__restrict__ h2e, int sorb) {
  const u32 h0 = e.h0 & 0xff, p0 = e.p0 & 0xff;
  if (!e.dbl) {
    T v = (T)0.0;
    v += __ldg(h1e + (size_t)p0 * sorb + h0);
#pragma unroll
    for (int i = 0; i < L; ++i) {
      u64 rest = x.w[i];
      while (rest) {
        const int b = 63 - __clzll((long long)rest);
        rest ^= 1ull << b;
        const u32 k = (u32)(64 * i + b);
        v += two_body<T>(h2e, h0, k, p0, k);
      }
    }
    const int sg = ((e.h0 ^ e.p0) >> 8) ^ (int)(h0 < p0);
    v *= (sg & 1) ? (T)-1.0 : (T)1.0;
    return v;
  }
  const u32 h1 = e.h1 & 0xff, p1 = e.p1 & 0xff;
  const u32 hh = h0 > h1 ? h0 : h1, hl = h0 > h1 ? h1 : h0;
  const u32 ph = p0 > p1 ? p0 : p1, pl = p0 > p1 ? p1 : p0;
  const int cross = (int)(hh < ph) + (int)(hl < ph) + (int)(hh < pl) + (int)(hl < pl);
  const int sg = ((e.h0 ^ e.h1 ^ e.p0 ^ e.p1) >> 8) ^ cross ^ 1;
  T v = (T)1.0 * __ldg(h2e + pair_offset(hh, hl, ph, pl));
  v *= (sg & 1) ? (T)-1.0 : (T)1.0;
  return v;
}

// diagonal element over the first `nele` occupied orbitals (padding with orbital 0 exactly as the
// reference's zero-initialised olst does)   cpp_src/cpu/hamiltonian.cpp:33-50
template <int L, typename T>
__device__ T diag_element(const Onv<L> &x, const T *__restrict__ h1e, const T *__restrict__ h2e, int sorb, int nele) {
  T v = (T)0.0;
  Onv<L> outer = x;
  for (int a = 0; a < nele; ++a) {
    u32 p = 0;
#pragma unroll
    for (int i = 0; i < L; ++i) {
      if (outer.w[i]) {
        const int b = __ffsll((long long)outer.w[i]) - 1;
        outer.w[i] ^= 1ull << b;
        p = (u32)(64 * i + b);
        break;
      }
    }
    v += __ldg(h1e + (size_t)p * sorb + p);
    Onv<L> inner = x;
    for (int c = 0; c < a; ++c) {
      u32 q = 0;
#pragma unroll
      for (int i = 0; i < L; ++i) {
        if (inner.w[i]) {
          const int b = __ffsll((long long)inner.w[i]) - 1;
          inner.w[i] ^= 1ull << b;
          q = (u32)(64 * i + b);
          break;
        }
      }
      v += two_body<T>(h2e, p, q, p, q);
    }
  }
  return v;
}

// compare little-endian multi-word integers: <0, 0, >0   (cpp_src/tensor/cpu_tensor.cpp:601-608)
template <int L>
__device__ __forceinline__ int cmp_onv(const Onv<L> &a, const Onv<L> &b) {
#pragma unroll
  for (int i = L - 1; i >= 0; --i) {
    if (a.w[i] < b.w[i]) return -1;
    if (a.w[i] > b.w[i]) return 1;
  }
  return 0;
}

template <int L>
__device__ __forceinline__ bool eq_onv(const Onv<L> &a, const Onv<L> &b) {
  bool e = true;
#pragma unroll
  for (int i = 0; i < L; ++i) e &= (a.w[i] == b.w[i]);
  return e;
}

// ---- host-side plumbing -------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int check_launch(const char *what);
void count_launch(int n = 1);

}  // namespace pynqs
