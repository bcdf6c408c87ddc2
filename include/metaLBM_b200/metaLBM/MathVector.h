// metaLBM/MathVector.h (B200 drop-in) -- the fixed-size vector type of the reference API
// (MathVector.h:11-389, StaticArray.h): brace-initialisable aggregate `MathVector<U, N> v = {{a, b, c}}`,
// `operator[]`, `data()`, `sum/dot/norm2` and component-wise arithmetic; `Position` = MathVector<unsigned, 3>.
// Only what host code of the step path touches is provided (no device code lives on this side of the C-ABI).
#pragma once

#include <cmath>
#include <ostream>

#include "Commons.h"

namespace lbm {

template <class U, unsigned int NumberComponents>
class MathVector {
 public:
  U sArray[NumberComponents > 0 ? NumberComponents : 1];

  constexpr const U& operator[](int i) const { return sArray[i]; }
  U& operator[](int i) { return sArray[i]; }
  U* data() { return sArray; }
  constexpr const U* data() const { return sArray; }
  static constexpr unsigned int size() { return NumberComponents; }

  U sum() const {
    U r = 0;
    for (unsigned int i = 0; i < NumberComponents; ++i) r += sArray[i];
    return r;
  }
  U dot(const MathVector& other) const {
    U r = sArray[0] * other[0];
    for (unsigned int i = 1; i < NumberComponents; ++i) r += sArray[i] * other[i];
    return r;
  }
  U norm2() const { return dot(*this); }

  MathVector& operator+=(const MathVector& o) { for (unsigned int i = 0; i < NumberComponents; ++i) sArray[i] += o[i]; return *this; }
  MathVector& operator-=(const MathVector& o) { for (unsigned int i = 0; i < NumberComponents; ++i) sArray[i] -= o[i]; return *this; }
  MathVector& operator*=(const U f) { for (unsigned int i = 0; i < NumberComponents; ++i) sArray[i] *= f; return *this; }
  MathVector& operator/=(const U f) { for (unsigned int i = 0; i < NumberComponents; ++i) sArray[i] /= f; return *this; }
};

template <class U, unsigned int N> MathVector<U, N> operator+(MathVector<U, N> a, const MathVector<U, N>& b) { return a += b; }
template <class U, unsigned int N> MathVector<U, N> operator-(MathVector<U, N> a, const MathVector<U, N>& b) { return a -= b; }
template <class U, unsigned int N> MathVector<U, N> operator*(const U f, MathVector<U, N> a) { return a *= f; }
template <class U, unsigned int N> MathVector<U, N> operator*(MathVector<U, N> a, const U f) { return a *= f; }
template <class U, unsigned int N> MathVector<U, N> operator/(MathVector<U, N> a, const U f) { return a /= f; }
template <class U, unsigned int N> bool operator==(const MathVector<U, N>& a, const MathVector<U, N>& b) {
  for (unsigned int i = 0; i < N; ++i) if (!(a[i] == b[i])) return false;
  return true;
}
template <class U, unsigned int N> std::ostream& operator<<(std::ostream& os, const MathVector<U, N>& v) {
  os << "[";
  for (unsigned int i = 0; i < N; ++i) os << (i ? " " : "") << v[i];
  return os << "]";
}

typedef MathVector<unsigned int, 3> Position;

}  // namespace lbm
