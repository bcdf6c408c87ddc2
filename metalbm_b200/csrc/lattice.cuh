// lattice.cuh -- compile-time lattice descriptors for the fused step kernels.
//
// Same tables and the same population ordering as the reference's Lattice<T, DdQq>
// (Lattice.h:80-143 D2Q5, :145-210 D2Q9, :460-532 D3Q15, :535-612 D3Q19, :614-703 D3Q27; the multi-speed lattices
// :213-288 D2Q13, :290-370 D2Q17, :372-458 D2Q21, :706-803 D3Q33 with halos of 2-3 nodes and their own sound speeds):
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

enum LatticeId { kD2Q5 = 0, kD2Q9 = 1, kD3Q15 = 2, kD3Q19 = 3, kD3Q27 = 4, kD2Q13 = 5, kD2Q17 = 6, kD2Q21 = 7, kD3Q33 = 8 };

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
// multi-speed lattices: weight by |c|^2 as well (Lattice.h:275-283, 355-364, 443-453, 786-796)
template <> MLBM_HD constexpr double weightByNorm<kD2Q13>(int n2) {
  return n2 == 0 ? 1.0 / 2.0 : (n2 == 1 ? 4.0 / 45.0 : (n2 == 2 ? 1.0 / 30.0 : 1.0 / 360.0));
}
template <> MLBM_HD constexpr double weightByNorm<kD2Q17>(int n2) {
  return n2 == 0 ? 0.121527777777777777777778
                 : (n2 == 2 ? 0.175781250000000000000000
                            : (n2 == 8 ? 0.014062500000000000000000 : (n2 == 9 ? 0.027777777777777777777778 : 0.001996527777777777777778)));
}
template <> MLBM_HD constexpr double weightByNorm<kD2Q21>(int n2) {
  return n2 == 0 ? 91. / 324.
                 : (n2 == 1 ? 1. / 12. : (n2 == 2 ? 2. / 27. : (n2 == 4 ? 7. / 360. : (n2 == 8 ? 1. / 432. : 1. / 1620.))));
}
template <> MLBM_HD constexpr double weightByNorm<kD3Q33>(int n2) {
  return n2 == 0 ? 0.177627658370520295649084
                 : (n2 == 1 ? 0.103315974899246818673111
                            : (n2 == 2 ? 0.000513472406731114352456 : (n2 == 3 ? 0.021333928148672240120078 : 0.004273899693974583187026)));
}
}  // namespace detail

#define MLBM_LATTICE_COMMON(ID, DIM, QQ, FACEQ) MLBM_LATTICE_GENERAL(ID, DIM, QQ, FACEQ, 1, 3.0)

#define MLBM_LATTICE_GENERAL(ID, DIM, QQ, FACEQ, HALO, INV_CS2)                              \
  static constexpr int id = ID;                                                              \
  static constexpr int D = DIM;                                                              \
  static constexpr int Q = QQ;                                                               \
  static constexpr int faceQ = FACEQ;                                                        \
  static constexpr int H = HALO; /* dimH: the largest |c| component */                       \
  static constexpr double inv_cs2 = INV_CS2;                                                 \
  MLBM_HD static constexpr int maxNorm2() {                                                  \
    int n = 0;                                                                               \
    for (int q = 0; q < QQ; ++q) n = norm2(q) > n ? norm2(q) : n;                            \
    return n;                                                                                \
  }                                                                                          \
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

// ---- multi-speed lattices: plain integer tables (components up to +-3) ----
#define MLBM_WIDE_TABLE(...)                                                     \
  MLBM_HD static constexpr int packed(int q, int d) {                            \
    constexpr int t[][3] = {__VA_ARGS__};                                        \
    return t[q][d];                                                              \
  }

template <> struct Lattice<kD2Q13> {
  MLBM_WIDE_TABLE({0, 0, 0}, {-1, 0, 0}, {-1, -1, 0}, {-1, 1, 0}, {-2, 0, 0}, {1, 0, 0}, {1, -1, 0}, {1, 1, 0}, {2, 0, 0},
                  {0, -1, 0}, {0, 1, 0}, {0, -2, 0}, {0, 2, 0})
  MLBM_LATTICE_GENERAL(kD2Q13, 2, 13, 4, 2, 3.0)
};

template <> struct Lattice<kD2Q17> {
  MLBM_WIDE_TABLE({0, 0, 0}, {-1, -1, 0}, {-1, 1, 0}, {-2, -2, 0}, {-2, 2, 0}, {-3, 0, 0}, {-3, -3, 0}, {-3, 3, 0}, {1, -1, 0},
                  {1, 1, 0}, {2, -2, 0}, {2, 2, 0}, {3, 0, 0}, {3, -3, 0}, {3, 3, 0}, {0, -3, 0}, {0, 3, 0})
  MLBM_LATTICE_GENERAL(kD2Q17, 2, 17, 7, 3, 2.0 / 3.0)   /* inv_cs2 = 2/3 as the reference has it (Lattice.h:294) */
};

template <> struct Lattice<kD2Q21> {
  MLBM_WIDE_TABLE({0, 0, 0}, {-1, 0, 0}, {-1, -1, 0}, {-1, 1, 0}, {-2, 0, 0}, {-2, 2, 0}, {-2, -2, 0}, {-3, 0, 0}, {1, 0, 0},
                  {1, -1, 0}, {1, 1, 0}, {2, 0, 0}, {2, -2, 0}, {2, 2, 0}, {3, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, -2, 0}, {0, 2, 0},
                  {0, -3, 0}, {0, 3, 0})
  MLBM_LATTICE_GENERAL(kD2Q21, 2, 21, 7, 3, 1.0 / (2.0 / 3.0))   /* cs2 = 2/3 (Lattice.h:377-378) */
};

template <> struct Lattice<kD3Q33> {
  MLBM_WIDE_TABLE({0, 0, 0}, {-1, 0, 0}, {-1, -1, 0}, {-1, 1, 0}, {-1, 0, -1}, {-1, 0, 1}, {-1, -1, -1}, {-1, -1, 1}, {-1, 1, -1},
                  {-1, 1, 1}, {-2, 0, 0}, {1, 0, 0}, {1, 1, 0}, {1, -1, 0}, {1, 0, 1}, {1, 0, -1}, {1, 1, 1}, {1, 1, -1}, {1, -1, 1},
                  {1, -1, -1}, {2, 0, 0}, {0, -1, 0}, {0, 0, -1}, {0, -1, -1}, {0, -1, 1}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1},
                  {0, 1, -1}, {0, 2, 0}, {0, -2, 0}, {0, 0, 2}, {0, 0, -2})
  MLBM_LATTICE_GENERAL(kD3Q33, 3, 33, 10, 2, 1.0 / 0.4156023517935171)   /* Lattice.h:710-711 */
};

inline int latticeDim(int id) { return (id <= kD2Q9 || (id >= kD2Q13 && id <= kD2Q21)) ? 2 : 3; }
inline int latticeHalo(int id) { return id == kD2Q13 || id == kD3Q33 ? 2 : (id == kD2Q17 || id == kD2Q21 ? 3 : 1); }
inline double latticeInvCs2(int id) {
  return id == kD2Q17 ? 2.0 / 3.0 : (id == kD2Q21 ? 1.0 / (2.0 / 3.0) : (id == kD3Q33 ? 1.0 / 0.4156023517935171 : 3.0));
}
inline int latticeQ(int id) {
  switch (id) {
    case kD2Q5: return 5;
    case kD2Q9: return 9;
    case kD3Q15: return 15;
    case kD3Q19: return 19;
    case kD3Q27: return 27;
    case kD2Q13: return 13;
    case kD2Q17: return 17;
    case kD2Q21: return 21;
    case kD3Q33: return 33;
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
    case kD2Q13: return 4;
    case kD2Q17: return 7;
    case kD2Q21: return 7;
    case kD3Q33: return 10;
    default: return 0;
  }
}

}  // namespace mlbm
