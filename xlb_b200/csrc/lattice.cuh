// Lattice tables as constexpr functions (usable with compile-time indices inside fully unrolled device loops, where
// they fold to immediates; no __constant__ traffic in the hot path).
//
// Index ORDER is the reference's (contractual):
//   D3Q27: l = 9*ix + 3*iy + iz with digit -> component {0 -> 0, 1 -> -1, 2 -> +1}   (xlb/velocity_set/d3q27.py:19,
//          itertools.product([0,-1,1], repeat=3))
//   D3Q19: the same enumeration with the 8 corners removed                             (d3q19.py:19)
//   D2Q9 : hand-listed                                                                 (d2q9.py:18-21)
// Derived tables follow xlb/velocity_set/velocity_set.py:124-221.
//
// Kernel coordinates vs physical coordinates: fields are [q][nx][ny][nz] with z unit-stride.  A 2-D field
// [q][nx][ny](1) is run as a 3-D array of extents (1, nx, ny) so that the unit-stride axis is still the thread axis:
// physical axis a maps to kernel axis a + (3 - D).  ck() are kernel-coordinate velocities, c() physical ones.
#pragma once

#include "common.cuh"

namespace xlbn {

#define XLBN_HD __host__ __device__ constexpr

struct D3Q27Base {
  static constexpr int D = 3, Q = 27, ID = XLBN_D3Q27;
  XLBN_HD static int digit(int i) { return i == 0 ? 0 : (i == 1 ? -1 : 1); }
  XLBN_HD static int ck(int axis, int l) { return axis == 0 ? digit(l / 9) : (axis == 1 ? digit((l / 3) % 3) : digit(l % 3)); }
  XLBN_HD static double w_by_speed(int s) { return s == 0 ? 8.0 / 27.0 : (s == 1 ? 2.0 / 27.0 : (s == 2 ? 1.0 / 54.0 : 1.0 / 216.0)); }
  XLBN_HD static int kaxis(int a) { return a; }  // kernel axis of physical axis a
};

struct D3Q19Base {
  static constexpr int D = 3, Q = 19, ID = XLBN_D3Q19;
  XLBN_HD static int idx27(int l) {
    constexpr int t[19] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 15, 18, 19, 20, 21, 24};
    return t[l];
  }
  XLBN_HD static int ck(int axis, int l) { return D3Q27Base::ck(axis, idx27(l)); }
  XLBN_HD static double w_by_speed(int s) { return s == 0 ? 1.0 / 3.0 : (s == 1 ? 1.0 / 18.0 : 1.0 / 36.0); }
  XLBN_HD static int kaxis(int a) { return a; }
};

struct D2Q9Base {
  static constexpr int D = 2, Q = 9, ID = XLBN_D2Q9;
  XLBN_HD static int ck(int axis, int l) {
    constexpr int cx[9] = {0, 0, 0, 1, -1, 1, -1, 1, -1};
    constexpr int cy[9] = {0, 1, -1, 0, 1, -1, 0, 1, -1};
    return axis == 0 ? 0 : (axis == 1 ? cx[l] : cy[l]);
  }
  XLBN_HD static double w_by_speed(int s) { return s == 0 ? 4.0 / 9.0 : (s == 1 ? 1.0 / 9.0 : 1.0 / 36.0); }
  XLBN_HD static int kaxis(int a) { return a + 1; }
};

// D2Q9 for x-slab runs: the same velocity set with physical x on kernel axis 0 (the slab axis, ghost planes) and physical y on the
// unit-stride thread axis; the kernel's middle axis has extent 1.  A [q][nx][ny] field is then run as kernel extents (nx, 1, ny).
struct D2Q9XBase {
  static constexpr int D = 2, Q = 9, ID = XLBN_D2Q9;
  XLBN_HD static int ck(int axis, int l) { return axis == 0 ? D2Q9Base::ck(1, l) : (axis == 1 ? 0 : D2Q9Base::ck(2, l)); }
  XLBN_HD static double w_by_speed(int s) { return D2Q9Base::w_by_speed(s); }
  XLBN_HD static int kaxis(int a) { return a == 0 ? 0 : 2; }
};

template <class B>
struct Lattice : B {
  using B::D;
  using B::Q;
  static constexpr int NT = D * (D + 1) / 2;  // independent components of a symmetric DxD tensor

  XLBN_HD static int iabs(int v) { return v < 0 ? -v : v; }
  XLBN_HD static int kaxis(int a) { return B::kaxis(a); }
  XLBN_HD static int c(int a, int l) { return B::ck(B::kaxis(a), l); }  // physical component a of velocity l
  XLBN_HD static int speed(int l) { return iabs(B::ck(0, l)) + iabs(B::ck(1, l)) + iabs(B::ck(2, l)); }
  XLBN_HD static double w(int l) { return B::w_by_speed(speed(l)); }
  XLBN_HD static int opp(int l) {
    for (int m = 0; m < Q; ++m)
      if (B::ck(0, m) == -B::ck(0, l) && B::ck(1, m) == -B::ck(1, l) && B::ck(2, m) == -B::ck(2, l)) return m;
    return -1;
  }
  XLBN_HD static bool is_main(int l) { return speed(l) == 1; }
  // t enumerates (a, b), a <= b, row-major: xx,xy,xz,yy,yz,zz (3-D) / xx,xy,yy (2-D)  (velocity_set.py:158-165)
  XLBN_HD static int pair_a(int t) {
    int k = 0;
    for (int a = 0; a < D; ++a)
      for (int b = a; b < D; ++b, ++k)
        if (k == t) return a;
    return -1;
  }
  XLBN_HD static int pair_b(int t) {
    int k = 0;
    for (int a = 0; a < D; ++a)
      for (int b = a; b < D; ++b, ++k)
        if (k == t) return b;
    return -1;
  }
  XLBN_HD static int cc(int l, int t) { return c(pair_a(t), l) * c(pair_b(t), l); }
  // Q_i = c c - cs^2 I with off-diagonals doubled (velocity_set.py:124-138)
  XLBN_HD static double qi(int l, int t) {
    return pair_a(t) == pair_b(t) ? (double)cc(l, t) - 1.0 / 3.0 : 2.0 * (double)cc(l, t);
  }
  // populations crossing an x-face: j-th population with ck(0) == sx, or -1  (velocity_set.py:195-221)
  XLBN_HD static int n_xdir() {
    int n = 0;
    for (int l = 0; l < Q; ++l) n += (B::ck(0, l) == 1);
    return n;
  }
  XLBN_HD static int xdir_slot(int l) {  // rank of l among the populations with the same non-zero ck(0)
    int n = 0;
    for (int m = 0; m < l; ++m) n += (B::ck(0, m) == B::ck(0, l));
    return n;
  }
};

using D3Q19 = Lattice<D3Q19Base>;
using D3Q27 = Lattice<D3Q27Base>;
using D2Q9 = Lattice<D2Q9Base>;
using D2Q9X = Lattice<D2Q9XBase>;

// Call fn(integral_constant<int, I>) for I = 0..N-1: a loop whose index is a constant expression in the body, so the
// lattice tables above are evaluated at compile time.
template <int I>
struct IC {
  static constexpr int value = I;
  __host__ __device__ constexpr operator int() const { return I; }
};
template <int N, int I = 0, class F>
XLBN_MATH void static_for(F&& fn) {
  if constexpr (I < N) {
    fn(IC<I>{});
    static_for<N, I + 1>(fn);
  }
}
template <int N, int I = 0, class F>
inline void static_for_host(F&& fn) {
  if constexpr (I < N) {
    fn(IC<I>{});
    static_for_host<N, I + 1>(fn);
  }
}

}  // namespace xlbn
