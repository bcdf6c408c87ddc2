// metaLBM/DynamicArray.h (B200 drop-in) -- host arrays with the reference's interface (DynamicArray.h:20-94,
// DynamicArray.cuh:88-127): data(), size(), operator[], copyFrom/copyTo.  CPUPinned is page-locked through the
// C-ABI (mlbm_alloc_pinned) so that Algorithm::pack/unpack and the stored fields move at PCIe speed; device
// arrays are owned by the mlbm_ctx and never appear on this side of the boundary.
#pragma once

#include <cstring>

#include "Commons.h"
#include "Options.h"

namespace lbm {

template <class U, Architecture architecture>
class DynamicArray {};

template <class U>
class DynamicArray<U, Architecture::CPU> {
 protected:
  unsigned int numberElements;
  U* dArrayPtr;

 public:
  DynamicArray(const unsigned int numberElements_in = 0)
      : numberElements(numberElements_in), dArrayPtr(numberElements_in ? static_cast<U*>(std::calloc(numberElements_in, sizeof(U))) : nullptr) {}
  DynamicArray(const DynamicArray& other) : DynamicArray(other.numberElements) { copyFrom(other); }
  DynamicArray& operator=(const DynamicArray&) = delete;
  virtual ~DynamicArray() { std::free(dArrayPtr); }

  U& operator[](int i) { return dArrayPtr[i]; }
  const U& operator[](int i) const { return dArrayPtr[i]; }
  U* data(const unsigned int offset = 0) { return dArrayPtr + offset; }
  const U* data(const unsigned int offset = 0) const { return dArrayPtr + offset; }
  unsigned int size() const { return numberElements; }
  template <class Other> void copyFrom(const Other& other) { std::memcpy(dArrayPtr, other.data(), sizeof(U) * other.size()); }
  template <class Other> void copyTo(Other& other) const { std::memcpy(other.data(), dArrayPtr, sizeof(U) * numberElements); }
};

template <class U>
class DynamicArray<U, Architecture::CPUPinned> {
 protected:
  unsigned int numberElements;
  U* dArrayPtr;

 public:
  DynamicArray(const unsigned int numberElements_in = 0) : numberElements(numberElements_in), dArrayPtr(nullptr) {
    if (numberElements) {
      void* pointer = nullptr;
      LBM_B200_CALL(mlbm_alloc_pinned(sizeof(U) * (size_t)numberElements, &pointer));
      dArrayPtr = static_cast<U*>(pointer);
      std::memset(dArrayPtr, 0, sizeof(U) * (size_t)numberElements);
    }
  }
  DynamicArray(const DynamicArray& other) : DynamicArray(other.numberElements) { copyFrom(other); }
  DynamicArray& operator=(const DynamicArray&) = delete;
  ~DynamicArray() { mlbm_free_pinned(dArrayPtr); }

  U& operator[](int i) { return dArrayPtr[i]; }
  const U& operator[](int i) const { return dArrayPtr[i]; }
  U* data(const unsigned int offset = 0) { return dArrayPtr + offset; }
  const U* data(const unsigned int offset = 0) const { return dArrayPtr + offset; }
  unsigned int size() const { return numberElements; }
  template <class Other> void copyFrom(const Other& other) { std::memcpy(dArrayPtr, other.data(), sizeof(U) * other.size()); }
  template <class Other> void copyTo(Other& other) const { std::memcpy(other.data(), dArrayPtr, sizeof(U) * numberElements); }
};

}  // namespace lbm
