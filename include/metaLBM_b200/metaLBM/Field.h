// metaLBM/Field.h (B200 drop-in) -- `Field<T, NumberComponents, architecture, IsWritten>` (Field.h:27-260): named
// SoA host arrays of `FFTWInit::numberElements` elements per component in the local padded layout.  As in the
// reference's GPU build the arrays of Architecture::GPU fields are page-locked HOST memory (Field.h:62-79): the
// reference lets the kernel write them over PCIe on stored steps, here mlbm_download_fields fills them after a
// stored step.  A field that is not written is not allocated and getData() returns NULL (Field.h:224-225).
#pragma once

#include <string>

#include "Commons.h"
#include "Computation.h"
#include "Domain.h"
#include "DynamicArray.h"
#include "FFTWInitializer.h"
#include "Lattice.h"
#include "MathVector.h"
#include "Options.h"
#include "Stream.h"

namespace lbm {

template <class T, unsigned int NumberComponents, Architecture architecture, bool IsWritten>
class Field {};

template <class T, unsigned int NumberComponents, Architecture architecture>
class Field<T, NumberComponents, architecture, true> {
 protected:
  DynamicArray<T, architecture == Architecture::GPU ? Architecture::CPUPinned : Architecture::CPU> array;

 public:
  static constexpr bool IsWritten = true;
  const std::string fieldName;

  Field(const std::string& fieldName_in) : array(FFTWInit::numberElements * NumberComponents), fieldName(fieldName_in) {}

  Field(const std::string& fieldName_in, const T& value_in, const Stream<architecture>&) : Field(fieldName_in) {
    for (unsigned int iC = 0; iC < NumberComponents; ++iC) fill(iC, value_in);
  }

  Field(const std::string& fieldName_in, const MathVector<T, NumberComponents>& vector_in, const Stream<architecture>&)
      : Field(fieldName_in) {
    for (unsigned int iC = 0; iC < NumberComponents; ++iC) fill(iC, vector_in[iC]);
  }

  Field(const Field& other) : array(other.array), fieldName(other.fieldName) {}

  T* getData(const unsigned int numberElements, const unsigned int iC = 0) { return array.data(iC * numberElements); }
  const T* getData(const unsigned int numberElements, const unsigned int iC = 0) const { return array.data(iC * numberElements); }
  DynamicArray<T, architecture == Architecture::GPU ? Architecture::CPUPinned : Architecture::CPU>& getArray() { return array; }

  void setValue(const Position& iP, const T value, const unsigned int numberElements, const unsigned int iC = 0) {
    getData(numberElements, iC)[lSD::getIndex(iP)] = value;
  }
  T getValue(const Position& iP, const unsigned int numberElements, const unsigned int iC = 0) const {
    return getData(numberElements, iC)[lSD::getIndex(iP)];
  }

 private:
  void fill(const unsigned int iC, const T value) {
    T* component = getData(FFTWInit::numberElements, iC);
    Computation<Architecture::CPU, L::dimD>(lSD::sStart(), lSD::sEnd()).Do([&](const Position& iP) { component[lSD::getIndex(iP)] = value; });
  }
};

template <class T, unsigned int NumberComponents, Architecture architecture>
class Field<T, NumberComponents, architecture, false> {
 public:
  static constexpr bool IsWritten = false;
  const std::string fieldName;
  Field(const std::string& fieldName_in) : fieldName(fieldName_in) {}
  Field(const std::string& fieldName_in, const T&, const Stream<architecture>&) : fieldName(fieldName_in) {}
  Field(const std::string& fieldName_in, const MathVector<T, NumberComponents>&, const Stream<architecture>&) : fieldName(fieldName_in) {}
  T* getData(const unsigned int, const unsigned int = 0) { return NULL; }
};

}  // namespace lbm
