// metaLBM/Lattice.h (B200 drop-in) -- `lbm::Lattice<T, LatticeType>` with the reference's member names
// (Lattice.h:19-806): inv_cs2, cs2, dimD, dimQ, dimH, faceQ, halo(), celerity(), weight(), iQ_Top/Bottom/Front/Back(),
// and the global aliases L / uiL (Lattice.h:804-806).
//
// The population ordering is part of the contract (SoA layout, halo planes, checkpoints): populations
// 1..faceQ have c_x < 0, faceQ+1..2*faceQ have c_x > 0 (Communication.h:138-176).  The tables below are the
// same numbers the CUDA kernels are specialised on (metalbm_b200/csrc/lattice.cuh); weights and the face lists
// are derived from the celerities instead of being spelled out a second time.
#pragma once

#include "Commons.h"
#include "MathVector.h"
#include "Options.h"

namespace lbm {

namespace b200 {

template <LatticeType> struct Stencil;  // dimD, dimQ, faceQ, abi (mlbm_lattice), c[dimQ][3], weightOfNorm2(n)

template <> struct Stencil<LatticeType::D2Q5> {
  static constexpr int dimD = 2, dimQ = 5, faceQ = 1, abi = MLBM_D2Q5;
  static constexpr int c(int q, int d) {
    constexpr int t[5][3] = {{0, 0, 0}, {-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}};
    return t[q][d];
  }
  static constexpr double weightOfNorm2(int n) { return n == 0 ? 4.0 / 6.0 : 1.0 / 12.0; }
};
template <> struct Stencil<LatticeType::D2Q9> {
  static constexpr int dimD = 2, dimQ = 9, faceQ = 3, abi = MLBM_D2Q9;
  static constexpr int c(int q, int d) {
    constexpr int t[9][3] = {{0, 0, 0}, {-1, 1, 0}, {-1, 0, 0}, {-1, -1, 0}, {1, -1, 0}, {1, 0, 0}, {1, 1, 0}, {0, -1, 0}, {0, 1, 0}};
    return t[q][d];
  }
  static constexpr double weightOfNorm2(int n) { return n == 0 ? 4.0 / 9.0 : (n == 1 ? 1.0 / 9.0 : 1.0 / 36.0); }
};
template <> struct Stencil<LatticeType::D3Q15> {
  static constexpr int dimD = 3, dimQ = 15, faceQ = 5, abi = MLBM_D3Q15;
  static constexpr int c(int q, int d) {
    constexpr int t[15][3] = {{0, 0, 0},  {-1, 0, 0}, {-1, -1, -1}, {-1, -1, 1}, {-1, 1, -1}, {-1, 1, 1}, {1, 0, 0}, {1, 1, 1},
                              {1, 1, -1}, {1, -1, 1}, {1, -1, -1},  {0, -1, 0},  {0, 0, -1},  {0, 1, 0},  {0, 0, 1}};
    return t[q][d];
  }
  static constexpr double weightOfNorm2(int n) { return n == 0 ? 2.0 / 9.0 : (n == 1 ? 1.0 / 9.0 : 1.0 / 72.0); }
};
template <> struct Stencil<LatticeType::D3Q19> {
  static constexpr int dimD = 3, dimQ = 19, faceQ = 5, abi = MLBM_D3Q19;
  static constexpr int c(int q, int d) {
    constexpr int t[19][3] = {{0, 0, 0},  {-1, 0, 0}, {-1, -1, 0}, {-1, 1, 0},  {-1, 0, -1}, {-1, 0, 1}, {1, 0, 0},
                              {1, 1, 0},  {1, -1, 0}, {1, 0, 1},   {1, 0, -1},  {0, -1, 0},  {0, 0, -1}, {0, -1, -1},
                              {0, -1, 1}, {0, 1, 0},  {0, 0, 1},   {0, 1, 1},   {0, 1, -1}};
    return t[q][d];
  }
  static constexpr double weightOfNorm2(int n) { return n == 0 ? 1.0 / 3.0 : (n == 1 ? 1.0 / 18.0 : 1.0 / 36.0); }
};
template <> struct Stencil<LatticeType::D3Q27> {
  static constexpr int dimD = 3, dimQ = 27, faceQ = 9, abi = MLBM_D3Q27;
  static constexpr int c(int q, int d) {
    constexpr int t[27][3] = {{0, 0, 0},   {-1, 0, 0},  {-1, -1, 0}, {-1, 1, 0},  {-1, 0, -1}, {-1, 0, 1}, {-1, -1, -1},
                              {-1, -1, 1}, {-1, 1, -1}, {-1, 1, 1},  {1, 0, 0},   {1, 1, 0},   {1, -1, 0}, {1, 0, 1},
                              {1, 0, -1},  {1, 1, 1},   {1, 1, -1},  {1, -1, 1},  {1, -1, -1}, {0, -1, 0}, {0, 0, -1},
                              {0, -1, -1}, {0, -1, 1},  {0, 1, 0},   {0, 0, 1},   {0, 1, 1},   {0, 1, -1}};
    return t[q][d];
  }
  static constexpr double weightOfNorm2(int n) {
    return n == 0 ? 8.0 / 27.0 : (n == 1 ? 2.0 / 27.0 : (n == 2 ? 1.0 / 54.0 : 1.0 / 216.0));
  }
};

// multi-speed lattices (Lattice.h:213-458, 706-803): jumps of up to dimH nodes, their own sound speeds; one GPU
template <> struct Stencil<LatticeType::D2Q13> {
  static constexpr int dimD = 2, dimQ = 13, faceQ = 4, abi = MLBM_D2Q13, dimH = 2;
  static constexpr double invCs2() { return 3.0; }
  static constexpr int c(int q, int d) {
    constexpr int t[13][3] = {{0, 0, 0}, {-1, 0, 0}, {-1, -1, 0}, {-1, 1, 0}, {-2, 0, 0}, {1, 0, 0}, {1, -1, 0},
                              {1, 1, 0}, {2, 0, 0},  {0, -1, 0},  {0, 1, 0},  {0, -2, 0}, {0, 2, 0}};
    return t[q][d];
  }
  static constexpr double weightOfNorm2(int n) { return n == 0 ? 1.0 / 2.0 : (n == 1 ? 4.0 / 45.0 : (n == 2 ? 1.0 / 30.0 : 1.0 / 360.0)); }
};
template <> struct Stencil<LatticeType::D2Q17> {
  static constexpr int dimD = 2, dimQ = 17, faceQ = 7, abi = MLBM_D2Q17, dimH = 3;
  static constexpr double invCs2() { return 2.0 / 3.0; }
  static constexpr int c(int q, int d) {
    constexpr int t[17][3] = {{0, 0, 0}, {-1, -1, 0}, {-1, 1, 0}, {-2, -2, 0}, {-2, 2, 0}, {-3, 0, 0}, {-3, -3, 0}, {-3, 3, 0}, {1, -1, 0},
                              {1, 1, 0}, {2, -2, 0},  {2, 2, 0},  {3, 0, 0},   {3, -3, 0}, {3, 3, 0},  {0, -3, 0},  {0, 3, 0}};
    return t[q][d];
  }
  static constexpr double weightOfNorm2(int n) {
    return n == 0 ? 0.121527777777777777777778
                  : (n == 2 ? 0.175781250000000000000000
                            : (n == 8 ? 0.014062500000000000000000 : (n == 9 ? 0.027777777777777777777778 : 0.001996527777777777777778)));
  }
};
template <> struct Stencil<LatticeType::D2Q21> {
  static constexpr int dimD = 2, dimQ = 21, faceQ = 7, abi = MLBM_D2Q21, dimH = 3;
  static constexpr double invCs2() { return 1.0 / (2.0 / 3.0); }
  static constexpr int c(int q, int d) {
    constexpr int t[21][3] = {{0, 0, 0}, {-1, 0, 0}, {-1, -1, 0}, {-1, 1, 0}, {-2, 0, 0}, {-2, 2, 0}, {-2, -2, 0}, {-3, 0, 0}, {1, 0, 0},
                              {1, -1, 0}, {1, 1, 0}, {2, 0, 0},   {2, -2, 0}, {2, 2, 0},  {3, 0, 0},  {0, -1, 0},  {0, 1, 0},  {0, -2, 0},
                              {0, 2, 0},  {0, -3, 0}, {0, 3, 0}};
    return t[q][d];
  }
  static constexpr double weightOfNorm2(int n) {
    return n == 0 ? 91. / 324. : (n == 1 ? 1. / 12. : (n == 2 ? 2. / 27. : (n == 4 ? 7. / 360. : (n == 8 ? 1. / 432. : 1. / 1620.))));
  }
};
template <> struct Stencil<LatticeType::D3Q33> {
  static constexpr int dimD = 3, dimQ = 33, faceQ = 10, abi = MLBM_D3Q33, dimH = 2;
  static constexpr double invCs2() { return 1.0 / 0.4156023517935171; }
  static constexpr int c(int q, int d) {
    constexpr int t[33][3] = {{0, 0, 0},   {-1, 0, 0},  {-1, -1, 0}, {-1, 1, 0}, {-1, 0, -1}, {-1, 0, 1}, {-1, -1, -1}, {-1, -1, 1}, {-1, 1, -1},
                              {-1, 1, 1},  {-2, 0, 0},  {1, 0, 0},   {1, 1, 0},  {1, -1, 0},  {1, 0, 1},  {1, 0, -1},   {1, 1, 1},   {1, 1, -1},
                              {1, -1, 1},  {1, -1, -1}, {2, 0, 0},   {0, -1, 0}, {0, 0, -1},  {0, -1, -1}, {0, -1, 1},  {0, 1, 0},   {0, 0, 1},
                              {0, 1, 1},   {0, 1, -1},  {0, 2, 0},   {0, -2, 0}, {0, 0, 2},   {0, 0, -2}};
    return t[q][d];
  }
  static constexpr double weightOfNorm2(int n) {
    return n == 0 ? 0.177627658370520295649084
                  : (n == 1 ? 0.103315974899246818673111
                            : (n == 2 ? 0.000513472406731114352456 : (n == 3 ? 0.021333928148672240120078 : 0.004273899693974583187026)));
  }
};

// dimH and inv_cs2 of a stencil: 1 and 3 unless it says otherwise
template <class S, class = void> struct StencilHalo { static constexpr int value = 1; static constexpr double invCs2() { return 3.0; } };
template <class S> struct StencilHalo<S, decltype((void)S::dimH)> { static constexpr int value = S::dimH; static constexpr double invCs2() { return S::invCs2(); } };

}  // namespace b200

template <class T, LatticeType LatticeT>
struct Lattice {
 private:
  using S = b200::Stencil<LatticeT>;
  // populations with celerity component `axis` equal to `sign`, ascending iQ (the order of Lattice.h:168-176, 558-576, 637-655)
  template <unsigned int Count>
  static constexpr MathVector<unsigned int, Count> face(int axis, int sign) {
    MathVector<unsigned int, Count> r = {};
    unsigned int n = 0;
    for (int q = 0; q < S::dimQ; ++q)
      if (axis < S::dimD && S::c(q, axis) * sign > 0 && n < Count) r.sArray[n++] = (unsigned int)q;
    return r;
  }

 public:
  static constexpr LatticeType Type = LatticeT;
  static constexpr int abi = S::abi;  // the mlbm_lattice value of the C-ABI

  static constexpr T inv_cs2 = (T)b200::StencilHalo<S>::invCs2();
  static constexpr T cs2 = (T)1 / inv_cs2;

  static constexpr int dimD = S::dimD;
  static constexpr int dimQ = S::dimQ;
  static constexpr int dimH = b200::StencilHalo<S>::value;
  static constexpr int faceQ = S::faceQ;

  static constexpr Position halo() { return Position{{dimH, dimD > 1 ? dimH : 0u, dimD > 2 ? dimH : 0u}}; }

  static constexpr MathVector<MathVector<T, dimD>, dimQ> celerity() {
    MathVector<MathVector<T, dimD>, dimQ> r = {};
    for (int q = 0; q < dimQ; ++q)
      for (int iD = 0; iD < dimD; ++iD) r.sArray[q].sArray[iD] = (T)S::c(q, iD);
    return r;
  }

  static constexpr MathVector<T, dimQ> weight() {
    MathVector<T, dimQ> r = {};
    for (int q = 0; q < dimQ; ++q) {
      int n = 0;
      for (int iD = 0; iD < dimD; ++iD) n += S::c(q, iD) * S::c(q, iD);
      r.sArray[q] = (T)S::weightOfNorm2(n);
    }
    return r;
  }

  static constexpr MathVector<unsigned int, (dimD > 1 ? faceQ : 0)> iQ_Bottom() { return face<(dimD > 1 ? faceQ : 0)>(1, -1); }
  static constexpr MathVector<unsigned int, (dimD > 1 ? faceQ : 0)> iQ_Top() { return face<(dimD > 1 ? faceQ : 0)>(1, 1); }
  static constexpr MathVector<unsigned int, (dimD > 2 ? faceQ : 0)> iQ_Front() { return face<(dimD > 2 ? faceQ : 0)>(2, -1); }
  static constexpr MathVector<unsigned int, (dimD > 2 ? faceQ : 0)> iQ_Back() { return face<(dimD > 2 ? faceQ : 0)>(2, 1); }
};

template <class T, LatticeType LatticeT> constexpr T Lattice<T, LatticeT>::inv_cs2;
template <class T, LatticeType LatticeT> constexpr T Lattice<T, LatticeT>::cs2;
template <class T, LatticeType LatticeT> constexpr int Lattice<T, LatticeT>::dimD;
template <class T, LatticeType LatticeT> constexpr int Lattice<T, LatticeT>::dimQ;
template <class T, LatticeType LatticeT> constexpr int Lattice<T, LatticeT>::dimH;
template <class T, LatticeType LatticeT> constexpr int Lattice<T, LatticeT>::faceQ;

typedef Lattice<dataT, latticeT> L;
typedef Lattice<unsigned int, latticeT> uiL;

}  // namespace lbm
