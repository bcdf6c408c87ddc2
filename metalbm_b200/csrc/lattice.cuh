// lattice.cuh -- compile-time lattice descriptors for the fused step kernels.
//
// Same tables and the same population ordering as the reference's Lattice<T, DdQq>
// (Lattice.h:80-143 D2Q5, :145-210 D2Q9, :460-532 D3Q15, :535-612 D3Q19, :614-703 D3Q27):
// the SoA layout, the halo-plane contract (iQ 1..faceQ have c_x < 0, faceQ+1..2*faceQ have
// c_x > 0) and every checkpoint written by the reference depend on that ordering.
//
// Axes.  The reference stores z fastest and x slowest in 3-D and y fastest in 2-D
// (Domain.h:88-91, 205-207).  The kernels use three generic axes instead:
//   x = slowest axis (the slab / halo axis), m = middle axis, r = unit-stride "row" axis
// 3-D: (x, m, r) = (x, y, z);  2-D: (x, m, r) = (x, -, y) with a middle extent of 1.
#pragma once

#include <cstdint>

#ifdef __CUDACC__
#define MLBM_HD __host__ __device__ __forceinline__
#else
#define MLBM_HD inline
#endif

namespace mlbm {

enum LatticeId { kD2Q5 = 0, kD2Q9 = 1, kD3Q15 = 2, kD3Q19 = 3, kD3Q27 = 4 };

template <int Id> struct Lattice;

namespace detail {
// weight by squared speed |c|^2 = 0, 1, 2, 3
template <int Id> MLBM_HD constexpr double weightByNorm(int n2);
template <> MLBM_HD constexpr double weightByNorm<kD2Q5>(int n2) { return n2 == 0 ? 4.0 / 6.0 : 1.0 / 12.0; }
template <> MLBM_HD constexpr double weightByNorm<kD2Q9>(int n2) {
  return n2 == 0 ? 4.0 / 9.0 : (n2 == 1 ? 1.0 / 9.0 : 1.0 / 36.0);
}
template <> MLBM_HD constexpr double weightByNorm<kD3Q15>(int n2) {
  return n2 == 0 ? 2.0 / 9.0 : (n2 == 1 ? 1.0 / 9.0 : 1.0 / 72.0);
}
template <> MLBM_HD constexpr double weightByNorm<kD3Q19>(int n2) {
  return n2 == 0 ? 1.0 / 3.0 : (n2 == 1 ? 1.0 / 18.0 : 1.0 / 36.0);
}
template <> MLBM_HD constexpr double weightByNorm<kD3Q27>(int n2) {
  return n2 == 0 ? 8.0 / 27.0 : (n2 == 1 ? 2.0 / 27.0 : (n2 == 2 ? 1.0 / 54.0 : 1.0 / 216.0));
}
}  // namespace detail

#define MLBM_LATTICE_COMMON(ID, DIM, QQ, FACEQ)                                              \
  static constexpr int id = ID;                                                              \
  static constexpr int D = DIM;                                                              \
  static constexpr int Q = QQ;                                                               \
  static constexpr int faceQ = FACEQ;                                                        \
  static constexpr int H = 1; /* dimH */                                                     \
  static constexpr double inv_cs2 = 3.0;                                                     \
  /* physical celerity component d (0 = x, 1 = y, 2 = z) */                                  \
  MLBM_HD static constexpr int c(int q, int d) { return d < D ? packed(q, d) : 0; }          \
  /* kernel axes: slab axis, middle axis, unit-stride axis */                                \
  MLBM_HD static constexpr int cx(int q) { return packed(q, 0); }                            \
  MLBM_HD static constexpr int cm(int q) { return D == 3 ? packed(q, 1) : 0; }               \
  MLBM_HD static constexpr int cr(int q) { return packed(q, D - 1); }                        \
  MLBM_HD static constexpr int norm2(int q) {                                                \
    return packed(q, 0) * packed(q, 0) + packed(q, 1) * packed(q, 1) +                       \
           (D == 3 ? packed(q, 2) * packed(q, 2) : 0);                                       \
  }                                                                                          \
  MLBM_HD static constexpr double w(int q) { return detail::weightByNorm<ID>(norm2(q)); }

// Each celerity is packed as 2 bits per component (0 -> 0, 1 -> +1, 3 -> -1) so that the table is a
// single integer constant per population that folds away after loop unrolling.
#define MLBM_C(a, b, c_) ((uint32_t)(((a) & 3) | (((b) & 3) << 2) | (((c_) & 3) << 4)))
#define MLBM_UNPACK(word, d) ((int)(((word) >> (2 * (d))) & 3u) == 3 ? -1 : (int)(((word) >> (2 * (d))) & 3u))

template <> struct Lattice<kD2Q5> {
  MLBM_HD static constexpr int packed(int q, int d) {
    constexpr uint32_t t[5] = {MLBM_C(0, 0, 0), MLBM_C(-1, 0, 0), MLBM_C(1, 0, 0), MLBM_C(0, -1, 0), MLBM_C(0, 1, 0)};
    return MLBM_UNPACK(t[q], d);
  }
  MLBM_LATTICE_COMMON(kD2Q5, 2, 5, 1)
};

template <> struct Lattice<kD2Q9> {
  MLBM_HD static constexpr int packed(int q, int d) {
    constexpr uint32_t t[9] = {MLBM_C(0, 0, 0),  MLBM_C(-1, 1, 0), MLBM_C(-1, 0, 0), MLBM_C(-1, -1, 0), MLBM_C(1, -1, 0),
                               MLBM_C(1, 0, 0),  MLBM_C(1, 1, 0),  MLBM_C(0, -1, 0), MLBM_C(0, 1, 0)};
    return MLBM_UNPACK(t[q], d);
  }
  MLBM_LATTICE_COMMON(kD2Q9, 2, 9, 3)
};

template <> struct Lattice<kD3Q15> {
  MLBM_HD static constexpr int packed(int q, int d) {
    constexpr uint32_t t[15] = {MLBM_C(0, 0, 0),   MLBM_C(-1, 0, 0), MLBM_C(-1, -1, -1), MLBM_C(-1, -1, 1), MLBM_C(-1, 1, -1),
                                MLBM_C(-1, 1, 1),  MLBM_C(1, 0, 0),  MLBM_C(1, 1, 1),    MLBM_C(1, 1, -1),  MLBM_C(1, -1, 1),
                                MLBM_C(1, -1, -1), MLBM_C(0, -1, 0), MLBM_C(0, 0, -1),   MLBM_C(0, 1, 0),   MLBM_C(0, 0, 1)};
    return MLBM_UNPACK(t[q], d);
  }
  MLBM_LATTICE_COMMON(kD3Q15, 3, 15, 5)
};

template <> struct Lattice<kD3Q19> {
  MLBM_HD static constexpr int packed(int q, int d) {
    constexpr uint32_t t[19] = {MLBM_C(0, 0, 0),  MLBM_C(-1, 0, 0), MLBM_C(-1, -1, 0), MLBM_C(-1, 1, 0),  MLBM_C(-1, 0, -1),
                                MLBM_C(-1, 0, 1), MLBM_C(1, 0, 0),  MLBM_C(1, 1, 0),   MLBM_C(1, -1, 0),  MLBM_C(1, 0, 1),
                                MLBM_C(1, 0, -1), MLBM_C(0, -1, 0), MLBM_C(0, 0, -1),  MLBM_C(0, -1, -1), MLBM_C(0, -1, 1),
                                MLBM_C(0, 1, 0),  MLBM_C(0, 0, 1),  MLBM_C(0, 1, 1),   MLBM_C(0, 1, -1)};
    return MLBM_UNPACK(t[q], d);
  }
  MLBM_LATTICE_COMMON(kD3Q19, 3, 19, 5)
};

template <> struct Lattice<kD3Q27> {
  MLBM_HD static constexpr int packed(int q, int d) {
    constexpr uint32_t t[27] = {
        MLBM_C(0, 0, 0),   MLBM_C(-1, 0, 0),  MLBM_C(-1, -1, 0),  MLBM_C(-1, 1, 0),  MLBM_C(-1, 0, -1), MLBM_C(-1, 0, 1),
        MLBM_C(-1, -1, -1), MLBM_C(-1, -1, 1), MLBM_C(-1, 1, -1), MLBM_C(-1, 1, 1),  MLBM_C(1, 0, 0),   MLBM_C(1, 1, 0),
        MLBM_C(1, -1, 0),  MLBM_C(1, 0, 1),   MLBM_C(1, 0, -1),   MLBM_C(1, 1, 1),   MLBM_C(1, 1, -1),  MLBM_C(1, -1, 1),
        MLBM_C(1, -1, -1), MLBM_C(0, -1, 0),  MLBM_C(0, 0, -1),   MLBM_C(0, -1, -1), MLBM_C(0, -1, 1),  MLBM_C(0, 1, 0),
        MLBM_C(0, 0, 1),   MLBM_C(0, 1, 1),   MLBM_C(0, 1, -1)};
    return MLBM_UNPACK(t[q], d);
  }
  MLBM_LATTICE_COMMON(kD3Q27, 3, 27, 9)
};

inline int latticeDim(int id) { return id <= kD2Q9 ? 2 : 3; }
inline int latticeQ(int id) {
  switch (id) {
    case kD2Q5: return 5;
    case kD2Q9: return 9;
    case kD3Q15: return 15;
    case kD3Q19: return 19;
    case kD3Q27: return 27;
    default: return 0;
  }
}
inline int latticeFaceQ(int id) {
  switch (id) {
    case kD2Q5: return 1;
    case kD2Q9: return 3;
    case kD3Q15: return 5;
    case kD3Q19: return 5;
    case kD3Q27: return 9;
    default: return 0;
  }
}

}  // namespace mlbm
