// prepare.cuh -- layout of the prepared-integral workspace (see prepare.cu).
#pragma once
#include "common.cuh"

namespace pynqs {

// element offsets of the four tables inside the workspace (pure function of sorb)
struct PrepLayout {
  long long off_ab, off_aa, off_bb, off_s, total;
  int na, npair;
};

inline PrepLayout prep_layout(int sorb) {
  PrepLayout p;
  const long long na = sorb / 2, npair = na * (na - 1) / 2;
  p.na = (int)na;
  p.npair = (int)npair;
  p.off_ab = 0;
  p.off_aa = na * na * na * na;
  p.off_bb = p.off_aa + npair * npair;
  p.off_s = p.off_bb + npair * npair;
  p.total = p.off_s + (long long)sorb * 2 * na * na;
  return p;
}

template <typename T>
struct PrepView {
  const T *ab, *aa, *bb, *s;
  int na, npair;
};

template <typename T>
inline PrepView<T> prep_view(const void *ws, int sorb) {
  const PrepLayout p = prep_layout(sorb);
  const T *b = reinterpret_cast<const T *>(ws);
  PrepView<T> v;
  v.ab = b + p.off_ab;
  v.aa = b + p.off_aa;
  v.bb = b + p.off_bb;
  v.s = b + p.off_s;
  v.na = p.na;
  v.npair = p.npair;
  return v;
}

}  // namespace pynqs
