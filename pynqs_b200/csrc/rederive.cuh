// rederive.cuh -- <x|H|y> from the two bit strings alone: the excitation is re-derived from x ^ y
// (diff_type cpu/onstate.cpp:10-20, diff_orb :34-55, dispatch cpu/hamiltonian.cpp:87-102) and then
// evaluated by the same exc_element / diag_element the fused operator uses, so both routes give
// bit-identical values.
#pragma once
#include "common.cuh"

namespace pynqs {

// highest set bit over all words, removing it
template <int L>
__device__ __forceinline__ int pop_highest(Onv<L> &d) {
#pragma unroll
  for (int i = L - 1; i >= 0; --i) {
    if (d.w[i]) {
      const int b = 63 - __clzll((long long)d.w[i]);
      d.w[i] ^= 1ull << b;
      return 64 * i + b;
    }
  }
  return 0;
}

template <int L, typename T>
__device__ __forceinline__ T rederived_element(const Onv<L> &x, const Onv<L> &y, const T *__restrict__ h1e,
                                               const T *__restrict__ h2e, int sorb, int nele) {
  Onv<L> cre, ann;  // bra-only / ket-only orbitals
  int nc = 0, na = 0;
#pragma unroll
  for (int i = 0; i < L; ++i) {
    const u64 d = x.w[i] ^ y.w[i];
    cre.w[i] = d & x.w[i];
    ann.w[i] = d & y.w[i];
    nc += __popcll(cre.w[i]);
    na += __popcll(ann.w[i]);
  }
  if (nc == 0 && na == 0) return diag_element<L, T>(x, h1e, h2e, sorb, nele);
  Exc e;
  if (nc == 1 && na == 1) {
    const int h = pop_highest<L>(cre), p = pop_highest<L>(ann);
    e.h0 = h | ((count_below<L>(x, h) & 1) << 8);
    e.p0 = p | ((count_below<L>(x, p) & 1) << 8);
    e.h1 = e.p1 = 0;
    e.dbl = false;
    return exc_element<L, T>(x, e, h1e, h2e, sorb);
  }
  if (nc == 2 && na == 2) {
    const int h0 = pop_highest<L>(cre), h1 = pop_highest<L>(cre);
    const int p0 = pop_highest<L>(ann), p1 = pop_highest<L>(ann);
    e.h0 = h0 | ((count_below<L>(x, h0) & 1) << 8);
    e.h1 = h1 | ((count_below<L>(x, h1) & 1) << 8);
    e.p0 = p0 | ((count_below<L>(x, p0) & 1) << 8);
    e.p1 = p1 | ((count_below<L>(x, p1) & 1) << 8);
    e.dbl = true;
    return exc_element<L, T>(x, e, h1e, h2e, sorb);
  }
  return (T)0.0;
}

}  // namespace pynqs
