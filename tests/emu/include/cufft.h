// tests/emu/include/cufft.h -- TEST INFRASTRUCTURE ONLY: the cuFFT entry points csrc/spectral.cu uses, as naive host
// transforms (O(n^2) per line: the emulated library only ever sees small grids).  Same conventions as cuFFT / FFTW:
// unnormalised, forward = exp(-2 pi i jk/n), D2Z keeps the first n/2 + 1 outputs of the last dimension.
#pragma once
#include <cmath>
#include <complex>
#include <vector>

typedef int cufftResult;
enum { CUFFT_SUCCESS = 0, CUFFT_INVALID_PLAN = 1, CUFFT_INVALID_VALUE = 4 };
typedef enum { CUFFT_D2Z = 0x6a, CUFFT_Z2Z = 0x69 } cufftType;
enum { CUFFT_FORWARD = -1, CUFFT_INVERSE = 1 };
typedef double cufftDoubleReal;
typedef double2 cufftDoubleComplex;

namespace cufft_emu {
struct Plan {
  int rank = 0;
  long long n[3] = {1, 1, 1};
  bool embedded = false;
  long long istride = 1, idist = 0, ostride = 1, odist = 0, batch = 0;
  cufftType type = CUFFT_D2Z;
};
inline std::vector<Plan*>& plans() { static auto* p = new std::vector<Plan*>(); return *p; }   // outlives every static destructor
inline Plan* get(int handle) { return handle >= 0 && handle < (int)plans().size() ? plans()[handle] : nullptr; }
// strided complex line transform, in place
inline void line(std::complex<double>* data, long long n, long long stride, int sign) {
  std::vector<std::complex<double>> out((size_t)n);
  for (long long k = 0; k < n; ++k) {
    std::complex<double> sum = 0.0;
    for (long long j = 0; j < n; ++j) sum += data[j * stride] * std::polar(1.0, sign * 2.0 * M_PI * (double)((j * k) % n) / (double)n);
    out[(size_t)k] = sum;
  }
  for (long long k = 0; k < n; ++k) data[k * stride] = out[(size_t)k];
}
}  // namespace cufft_emu

typedef int cufftHandle;
inline cufftResult cufftCreate(cufftHandle* handle) {
  cufft_emu::plans().push_back(new cufft_emu::Plan());
  *handle = (int)cufft_emu::plans().size() - 1;
  return CUFFT_SUCCESS;
}
inline cufftResult cufftMakePlanMany64(cufftHandle handle, int rank, long long* n, long long* inembed, long long istride, long long idist,
                                       long long* onembed, long long ostride, long long odist, cufftType type, long long batch, size_t* workSize) {
  cufft_emu::Plan* p = cufft_emu::get(handle);
  if (!p || rank < 1 || rank > 2) return CUFFT_INVALID_VALUE;
  p->rank = rank;
  for (int i = 0; i < rank; ++i) p->n[i] = n[i];
  p->embedded = inembed != nullptr && onembed != nullptr;
  if (p->embedded && rank != 1) return CUFFT_INVALID_VALUE;   // the advanced layout is only used for the 1-D transform along x
  p->istride = istride; p->idist = idist; p->ostride = ostride; p->odist = odist;
  p->type = type;
  p->batch = batch;
  if (workSize) *workSize = 0;
  return CUFFT_SUCCESS;
}
inline cufftResult cufftSetStream(cufftHandle handle, cudaStream_t) { return cufft_emu::get(handle) ? CUFFT_SUCCESS : CUFFT_INVALID_PLAN; }
inline cufftResult cufftDestroy(cufftHandle handle) {
  cufft_emu::Plan* p = cufft_emu::get(handle);
  if (!p) return CUFFT_INVALID_PLAN;
  delete p;
  cufft_emu::plans()[handle] = nullptr;
  return CUFFT_SUCCESS;
}
// contiguous batches: real [batch][n0][n1] -> complex [batch][n0][n1/2 + 1] (rank 1: n0 == 1)
inline cufftResult cufftExecD2Z(cufftHandle handle, cufftDoubleReal* in, cufftDoubleComplex* out) {
  cufft_emu::Plan* p = cufft_emu::get(handle);
  if (!p || p->type != CUFFT_D2Z || p->embedded) return CUFFT_INVALID_PLAN;
  const long long n0 = p->rank == 2 ? p->n[0] : 1, n1 = p->rank == 2 ? p->n[1] : p->n[0], half = n1 / 2 + 1;
  std::vector<std::complex<double>> plane((size_t)(n0 * n1));
  for (long long b = 0; b < p->batch; ++b) {
    for (long long i = 0; i < n0 * n1; ++i) plane[(size_t)i] = in[b * n0 * n1 + i];
    for (long long row = 0; row < n0; ++row) cufft_emu::line(plane.data() + row * n1, n1, 1, -1);
    if (n0 > 1) for (long long column = 0; column < half; ++column) cufft_emu::line(plane.data() + column, n0, n1, -1);
    for (long long row = 0; row < n0; ++row)
      for (long long column = 0; column < half; ++column) {
        const std::complex<double> v = plane[(size_t)(row * n1 + column)];
        out[b * n0 * half + row * half + column] = double2{v.real(), v.imag()};
      }
  }
  return CUFFT_SUCCESS;
}
// 1-D complex transforms with the advanced layout: element j of batch b at data[b * dist + j * stride]
inline cufftResult cufftExecZ2Z(cufftHandle handle, cufftDoubleComplex* in, cufftDoubleComplex* out, int direction) {
  cufft_emu::Plan* p = cufft_emu::get(handle);
  if (!p || p->type != CUFFT_Z2Z || p->rank != 1 || in != out) return CUFFT_INVALID_PLAN;
  const long long stride = p->embedded ? p->istride : 1, dist = p->embedded ? p->idist : p->n[0];
  static_assert(sizeof(std::complex<double>) == sizeof(double2), "layout");
  for (long long b = 0; b < p->batch; ++b) cufft_emu::line(reinterpret_cast<std::complex<double>*>(in) + b * dist, p->n[0], stride, direction);
  return CUFFT_SUCCESS;
}
