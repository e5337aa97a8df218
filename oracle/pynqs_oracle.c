/*
 * oracle/pynqs_oracle.c -- CPU restatement of the PyNQS local-energy arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in pynqs_b200/ may include, link or call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * use it, as the checker.  Parity status: PINNED -- tests/test_oracle_golden.py checks every
 * function here bit-for-bit against the unmodified reference extension (oracle/_ref, built by
 * oracle/build_ref.py) via the fixtures in tests/golden/, and against the docstring
 * known-answer examples of libs/C_extension.pyi.
 *
 * Plain C99, scalar; one function per reference routine, each citing the file:line under
 * /root/reference/ whose behaviour it restates.  ONVs are little-endian multi-word bit strings:
 * spin orbital s lives in bit (s % 64) of word (s / 64); even = alpha, odd = beta.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;

static inline int words_for(int sorb) { return (sorb - 1) / 64 + 1; }
static inline int bit_of(const u64 *v, int k) { return (int)((v[k >> 6] >> (k & 63)) & 1ULL); }
static inline void flip(u64 *v, int k) { v[k >> 6] ^= 1ULL << (k & 63); }

/* parity of the occupation below orbital n: +1 / -1.   cpp_src/cpu/onstate.cpp:22-32 */
static int sign_below(const u64 *v, int n) {
  int par = 0;
  for (int w = 0; w < n / 64; ++w) par ^= __builtin_parityll(v[w]);
  if (n % 64) par ^= __builtin_parityll(v[n / 64] & ((1ULL << (n % 64)) - 1ULL));
  return 1 - 2 * par;
}

/* number of singles+doubles.   cpp_src/cpu/excitation.cpp:8-16 */
int orc_num_sd(int sorb, int noA, int noB) {
  int k = sorb / 2, nvA = k - noA, nvB = k - noB;
  return noA * nvA + noB * nvB + noA * (noA - 1) * nvA * (nvA - 1) / 4 +
         noB * (noB - 1) * nvB * (nvB - 1) / 4 + noA * noB * nvA * nvB;
}

/* merged alpha/beta interleaved list: occupied ascending, then virtual ascending; slot 2t is the
 * t-th alpha entry, slot 2t+1 the t-th beta entry.   cpp_src/cpu/onstate.cpp:147-193 */
static void merged_list(const u64 *bra, int sorb, int *lst) {
  int na = 0, nb = 0;
  for (int pass = 0; pass < 2; ++pass)
    for (int s = 0; s < sorb; ++s)
      if (bit_of(bra, s) == (pass == 0)) {
        if (s & 1) lst[2 * (nb++) + 1] = s;
        else lst[2 * (na++)] = s;
      }
}

void orc_merged(const u64 *bra, int64_t n, int sorb, int32_t *merged) {
  int L = words_for(sorb);
  for (int64_t s = 0; s < n; ++s) merged_list(bra + s * L, sorb, merged + s * sorb);
}

/* triangular pair index t -> (i, j), i > j >= 0.   cpp_src/cpu/excitation.h:6-11 */
static void tri_unpack(int t, int *i, int *j) {
  int a = (int)(sqrt((double)((t + 1) * 2)) + 0.5);
  *i = a;
  *j = t - a * (a - 1) / 2;
}

/* flat excitation index r -> four slots of the merged list + single(0)/double(1) flag.
 * Block order S_alpha, S_beta, D_aa, D_bb, D_ab; same-spin doubles take the hole pair from the
 * GLOBAL index r (r % noAA), not the block-local one.   cpp_src/cpu/excitation.cpp:18-110 */
void orc_unpack(int sorb, int noA, int noB, int r, int *out) {
  int k = sorb / 2, nvA = k - noA, nvB = k - noB;
  int noAA = noA * (noA - 1) / 2, noBB = noB * (noB - 1) / 2;
  int nvAA = nvA * (nvA - 1) / 2, nvBB = nvB * (nvB - 1) / 2;
  int d0 = noA * nvA, d1 = d0 + noB * nvB, d2 = d1 + noAA * nvAA, d3 = d2 + noBB * nvBB;
  int i = -1, a = -1, j = -1, b = -1, dbl = 0;
  if (r < d0) {
    i = 2 * (r % noA);
    a = 2 * (r / noA + noA);
    j = b = 0;
  } else if (r < d1) {
    int q = r - d0;
    i = 2 * (q % noB) + 1;
    a = 2 * (q / noB + noB) + 1;
    j = b = 0;
  } else if (r < d2) {
    int q = r - d1, hi, lo, vh, vl;
    tri_unpack(r % noAA, &hi, &lo);
    tri_unpack(q / noAA, &vh, &vl);
    i = 2 * hi; j = 2 * lo; a = 2 * (vh + noA); b = 2 * (vl + noA);
    dbl = 1;
  } else if (r < d3) {
    int q = r - d2, hi, lo, vh, vl;
    tri_unpack(r % noBB, &hi, &lo);
    tri_unpack(q / noBB, &vh, &vl);
    i = 2 * hi + 1; j = 2 * lo + 1; a = 2 * (vh + noB) + 1; b = 2 * (vl + noB) + 1;
    dbl = 1;
  } else {
    int q = r - d3, ia = q % (noA * nvA), jb = q / (noA * nvA);
    i = 2 * (ia % noA);
    a = 2 * (ia / noA + noA);
    j = 2 * (jb % noB) + 1;
    b = 2 * (jb / noB + noB) + 1;
    dbl = 1;
  }
  out[0] = i; out[1] = a; out[2] = j; out[3] = b; out[4] = dbl;
}

/* all connected determinants: row 0 = bra, row r+1 = bra with the four merged-list slots of r
 * flipped.   cpp_src/cpu/excitation.cpp:112-122, cpp_src/tensor/cpu_tensor.cpp:164-218 */
void orc_comb(const u64 *bra, int64_t n, int sorb, int noA, int noB, u64 *comb) {
  int L = words_for(sorb);
  int64_t M = (int64_t)orc_num_sd(sorb, noA, noB) + 1;
  int *lst = (int *)malloc(sizeof(int) * (size_t)sorb);
  for (int64_t s = 0; s < n; ++s) {
    const u64 *x = bra + s * L;
    merged_list(x, sorb, lst);
    for (int64_t m = 0; m < M; ++m) {
      u64 *row = comb + (s * M + m) * L;
      memcpy(row, x, sizeof(u64) * (size_t)L);
      if (m == 0) continue;
      int sl[5];
      orc_unpack(sorb, noA, noB, (int)(m - 1), sl);
      for (int t = 0; t < 4; ++t) flip(row, lst[sl[t]]);
    }
  }
  free(lst);
}

/* "excitation indices" and "signs" of the parity contract: for every row r+1 the four orbitals
 * (hole, particle, hole2, particle2; for singles hole2 = particle2 = merged[0], a no-op pair)
 * and the fermionic sign used by the fused routine.   cpp_src/cpu/excitation.cpp:124-169 */
void orc_excitations(const u64 *bra, int64_t n, int sorb, int noA, int noB, int32_t *orbs, int8_t *sgn) {
  int L = words_for(sorb);
  int64_t nsd = orc_num_sd(sorb, noA, noB);
  int *lst = (int *)malloc(sizeof(int) * (size_t)sorb);
  u64 ket[4];
  for (int64_t s = 0; s < n; ++s) {
    const u64 *x = bra + s * L;
    merged_list(x, sorb, lst);
    for (int64_t r = 0; r < nsd; ++r) {
      int sl[5], o[4];
      orc_unpack(sorb, noA, noB, (int)r, sl);
      memcpy(ket, x, sizeof(u64) * (size_t)L);
      for (int t = 0; t < 4; ++t) { o[t] = lst[sl[t]]; flip(ket, o[t]); }
      int sg;
      if (!sl[4]) sg = sign_below(x, o[0]) * sign_below(ket, o[1]);
      else {
        int p0 = o[0] > o[2] ? o[0] : o[2], p1 = o[0] > o[2] ? o[2] : o[0];
        int q0 = o[1] > o[3] ? o[1] : o[3], q1 = o[1] > o[3] ? o[3] : o[1];
        sg = sign_below(x, p0) * sign_below(x, p1) * sign_below(ket, q0) * sign_below(ket, q1);
      }
      for (int t = 0; t < 4; ++t) orbs[(s * nsd + r) * 4 + t] = o[t];
      sgn[s * nsd + r] = (int8_t)sg;
    }
  }
  free(lst);
}

/* ---- arithmetic in two precisions (the reference dispatches float32 / float64) ------------ */
#define DEFINE_PRECISION(T, SFX)                                                                   \
  /* h1e[j*sorb + i].   cpp_src/cpu/hamiltonian.cpp:7-11 */                                        \
  static T one_body_##SFX(const T *h1e, size_t i, size_t j, size_t sorb) { return h1e[j * sorb + i]; } \
  /* antisymmetrised <ij||kl> from the packed array.   cpp_src/cpu/hamiltonian.cpp:13-31 */        \
  static T two_body_##SFX(const T *h2e, size_t i, size_t j, size_t k, size_t l) {                  \
    if (i == j || k == l) return (T)0.0;                                                           \
    size_t ij = i > j ? i * (i - 1) / 2 + j : j * (j - 1) / 2 + i;                                  \
    size_t kl = k > l ? k * (k - 1) / 2 + l : l * (l - 1) / 2 + k;                                  \
    T s = (T)1;                                                                                    \
    if (!(i > j)) s = -s;                                                                          \
    if (!(k > l)) s = -s;                                                                          \
    size_t off = ij >= kl ? ij * (ij + 1) / 2 + kl : kl * (kl + 1) / 2 + ij;                        \
    return s * h2e[off];                                                                           \
  }                                                                                                \
  /* diagonal: p ascending, q < p ascending.   cpp_src/cpu/hamiltonian.cpp:33-50 */                \
  static T diag_##SFX(const u64 *x, const T *h1e, const T *h2e, int sorb, int nele, int L) {       \
    int occ[64 * 3];                                                                               \
    int no = 0;                                                                                    \
    memset(occ, 0, sizeof(occ));                                                                   \
    for (int s = 0; s < 64 * L; ++s)                                                               \
      if (bit_of(x, s)) occ[no++] = s;                                                             \
    T v = (T)0.0;                                                                                  \
    for (int a = 0; a < nele; ++a) {                                                               \
      int p = occ[a];                                                                              \
      v += one_body_##SFX(h1e, (size_t)p, (size_t)p, (size_t)sorb);                                \
      for (int b = 0; b < a; ++b) v += two_body_##SFX(h2e, (size_t)p, (size_t)occ[b], (size_t)p, (size_t)occ[b]); \
    }                                                                                              \
    return v;                                                                                      \
  }                                                                                                \
  /* single p->q: h1e + sum over occupied k, words ascending, bits DEscending.                     \
   * cpp_src/cpu/hamiltonian.cpp:52-72, cpp_src/cpu/excitation.cpp:141-156 */                      \
  static T single_##SFX(const u64 *x, const u64 *ket, int p, int q, const T *h1e, const T *h2e,    \
                        int sorb, int L) {                                                         \
    T v = (T)0.0;                                                                                  \
    v += one_body_##SFX(h1e, (size_t)p, (size_t)q, (size_t)sorb);                                  \
    for (int w = 0; w < L; ++w)                                                                    \
      for (int b = 63; b >= 0; --b)                                                                \
        if ((x[w] >> b) & 1ULL) {                                                                  \
          size_t k = (size_t)(64 * w + b);                                                         \
          v += two_body_##SFX(h2e, (size_t)p, k, (size_t)q, k);                                    \
        }                                                                                          \
    v *= (T)(sign_below(x, p) * sign_below(ket, q));                                               \
    return v;                                                                                      \
  }                                                                                                \
  /* double p0>p1 -> q0>q1.   cpp_src/cpu/hamiltonian.cpp:74-85, excitation.cpp:157-167 */         \
  static T double_##SFX(const u64 *x, const u64 *ket, int p0, int p1, int q0, int q1, const T *h2e) { \
    int sg = sign_below(x, p0) * sign_below(x, p1) * sign_below(ket, q0) * sign_below(ket, q1);    \
    T v = two_body_##SFX(h2e, (size_t)p0, (size_t)p1, (size_t)q0, (size_t)q1);                     \
    v *= (T)sg;                                                                                    \
    return v;                                                                                      \
  }                                                                                                \
  /* fused enumerate + H_ij.  Row 0 = <x|H|x>.   cpp_src/tensor/cpu_tensor.cpp:220-272 */          \
  void orc_comb_hij_##SFX(const u64 *bra, const T *h1e, const T *h2e, int64_t n, int sorb, int nele, \
                          int noA, int noB, u64 *comb, T *hmat) {                                  \
    int L = words_for(sorb);                                                                       \
    int64_t M = (int64_t)orc_num_sd(sorb, noA, noB) + 1;                                           \
    int *lst = (int *)malloc(sizeof(int) * (size_t)sorb);                                          \
    for (int64_t s = 0; s < n; ++s) {                                                              \
      const u64 *x = bra + s * L;                                                                  \
      merged_list(x, sorb, lst);                                                                   \
      memcpy(comb + s * M * L, x, sizeof(u64) * (size_t)L);                                        \
      hmat[s * M] = diag_##SFX(x, h1e, h2e, sorb, nele, L);                                        \
      for (int64_t m = 1; m < M; ++m) {                                                            \
        u64 *row = comb + (s * M + m) * L;                                                         \
        memcpy(row, x, sizeof(u64) * (size_t)L);                                                   \
        int sl[5], o[4];                                                                           \
        orc_unpack(sorb, noA, noB, (int)(m - 1), sl);                                              \
        for (int t = 0; t < 4; ++t) { o[t] = lst[sl[t]]; flip(row, o[t]); }                        \
        if (!sl[4]) hmat[s * M + m] = single_##SFX(x, row, o[0], o[1], h1e, h2e, sorb, L);         \
        else {                                                                                     \
          int p0 = o[0] > o[2] ? o[0] : o[2], p1 = o[0] > o[2] ? o[2] : o[0];                      \
          int q0 = o[1] > o[3] ? o[1] : o[3], q1 = o[1] > o[3] ? o[3] : o[1];                      \
          hmat[s * M + m] = double_##SFX(x, row, p0, p1, q0, q1, h2e);                             \
        }                                                                                          \
      }                                                                                            \
    }                                                                                              \
    free(lst);                                                                                     \
  }                                                                                                \
  /* <bra|H|ket> re-deriving the excitation from the two bit strings; >2-fold -> 0.                \
   * cpp_src/cpu/hamiltonian.cpp:87-102; diff_type onstate.cpp:10-20; diff_orb onstate.cpp:34-55 */ \
  static T element_##SFX(const u64 *x, const u64 *y, const T *h1e, const T *h2e, int sorb, int nele, int L) { \
    int nc = 0, na = 0, cre[2], ann[2];                                                            \
    for (int w = L - 1; w >= 0; --w) {                                                             \
      u64 d = x[w] ^ y[w];                                                                         \
      nc += __builtin_popcountll(d & x[w]);                                                        \
      na += __builtin_popcountll(d & y[w]);                                                        \
    }                                                                                              \
    if (nc == 0 && na == 0) return diag_##SFX(x, h1e, h2e, sorb, nele, L);                         \
    if (!((nc == 1 && na == 1) || (nc == 2 && na == 2))) return (T)0.0;                            \
    int ic = 0, ia = 0;                                                                            \
    for (int w = L - 1; w >= 0; --w)                                                               \
      for (int b = 63; b >= 0; --b) {                                                              \
        u64 m = 1ULL << b, d = x[w] ^ y[w];                                                        \
        if (d & x[w] & m) cre[ic++] = 64 * w + b;                                                  \
        if (d & y[w] & m) ann[ia++] = 64 * w + b;                                                  \
      }                                                                                            \
    if (nc == 1) return single_##SFX(x, y, cre[0], ann[0], h1e, h2e, sorb, L);                     \
    return double_##SFX(x, y, cre[0], cre[1], ann[0], ann[1], h2e);                                \
  }                                                                                                \
  /* Hmat[n,m]: ket3d ? ket[n,m,L] : ket[m,L].   cpp_src/tensor/cpu_tensor.cpp:274-325 */          \
  void orc_hij_##SFX(const u64 *bra, const u64 *ket, const T *h1e, const T *h2e, int64_t n, int64_t m, \
                     int ket3d, int sorb, int nele, T *out) {                                      \
    int L = words_for(sorb);                                                                       \
    for (int64_t i = 0; i < n; ++i)                                                                \
      for (int64_t j = 0; j < m; ++j) {                                                            \
        const u64 *y = ket + ((ket3d ? i * m : 0) + j) * L;                                        \
        out[i * m + j] = element_##SFX(bra + i * L, y, h1e, h2e, sorb, nele, L);                   \
      }                                                                                            \
  }                                                                                                \
  /* +1 occupied / -1 empty.   cpp_src/cpu/onstate.h:44-63, cpu_tensor.cpp:46-88 */                \
  void orc_onv_to_tensor_##SFX(const u64 *bra, int64_t n, int sorb, T *out) {                      \
    int L = words_for(sorb);                                                                       \
    for (int64_t s = 0; s < n; ++s)                                                                \
      for (int k = 0; k < sorb; ++k) out[s * sorb + k] = bit_of(bra + s * L, k) ? (T)1.0 : (T)-1.0; \
  }

DEFINE_PRECISION(double, f64)
DEFINE_PRECISION(float, f32)

/* 0/1 bytes -> packed words (only the value 1 sets a bit).   cpp_src/tensor/cpu_tensor.cpp:8-44 */
void orc_tensor_to_onv(const uint8_t *states, int64_t n, int sorb, u64 *out) {
  int L = words_for(sorb);
  memset(out, 0, sizeof(u64) * (size_t)(n * L));
  for (int64_t s = 0; s < n; ++s)
    for (int k = 0; k < sorb; ++k)
      if (states[s * sorb + k] == 1) flip(out + s * L, k);
}

/* classic binary search over rows sorted as little-endian multi-word integers (word L-1 most
 * significant); -1 / false when absent.   cpp_src/tensor/cpu_tensor.cpp:589-688 */
void orc_lut(const u64 *key, int64_t N, const u64 *q, int64_t n, int L, int64_t *idx, uint8_t *mask) {
  for (int64_t t = 0; t < n; ++t) {
    const u64 *x = q + t * L;
    int64_t lo = 0, hi = N - 1, found = -1;
    while (lo <= hi) {
      int64_t mid = lo + (hi - lo) / 2;
      const u64 *e = key + mid * L;
      int c = 0;
      for (int w = L - 1; w >= 0; --w) {
        if (e[w] < x[w]) { c = -1; break; }
        if (e[w] > x[w]) { c = 1; break; }
      }
      if (c == 0) { found = mid; break; }
      if (c < 0) lo = mid + 1; else hi = mid - 1;
    }
    idx[t] = found;
    mask[t] = found >= 0;
  }
}

/* sample-space local energy from materialised rows: psi'[m] = table value or 0,
 * E_loc = sum_m (psi'[m] / psi'[0]) * H[m] (complex psi as re/im pairs; cpsi = 1).
 * vmc/energy/eloc.py:383-397 with WavefunctionLUT.lookup utils/public_function.py:817-838.
 * The division is done per element, as in the reference; the sum runs left to right. */
void orc_eloc_rows(const int64_t *idx, const double *hmat, const double *psi, int cpsi, int64_t n, int64_t M,
                   double *eloc) {
  for (int64_t s = 0; s < n; ++s) {
    const int64_t *id = idx + s * M;
    double er = 0.0, ei = 0.0;
    if (!cpsi) {
      double p0 = id[0] >= 0 ? psi[id[0]] : 0.0;
      for (int64_t m = 0; m < M; ++m) {
        double pm = id[m] >= 0 ? psi[id[m]] : 0.0;
        er += (pm / p0) * hmat[s * M + m];
      }
      eloc[s] = er;
    } else {
      double c = id[0] >= 0 ? psi[2 * id[0]] : 0.0, d = id[0] >= 0 ? psi[2 * id[0] + 1] : 0.0;
      for (int64_t m = 0; m < M; ++m) {
        double a = id[m] >= 0 ? psi[2 * id[m]] : 0.0, b = id[m] >= 0 ? psi[2 * id[m] + 1] : 0.0;
        double qr, qi; /* numpy-style complex division, as torch's c10::complex operator/= */
        if (fabs(c) >= fabs(d)) {
          if (c == 0.0 && d == 0.0) { qr = a / fabs(c); qi = b / fabs(d); }
          else { double rat = d / c, scl = 1.0 / (c + d * rat); qr = (a + b * rat) * scl; qi = (b - a * rat) * scl; }
        } else { double rat = c / d, scl = 1.0 / (d + c * rat); qr = (a * rat + b) * scl; qi = (b * rat - a) * scl; }
        er += qr * hmat[s * M + m];
        ei += qi * hmat[s * M + m];
      }
      eloc[2 * s] = er;
      eloc[2 * s + 1] = ei;
    }
  }
}
