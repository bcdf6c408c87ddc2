/* oracle/shim/fftw3-mpi.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Stand-in for FFTW3-MPI (not vendored by the reference, no version pinned:
 * cmake/FindFFTW.cmake) so that the reference's CPU code compiles here.  Call
 * sites: FFTWInitializer.h:18-33, Transformer.h:37-60,81-104, DynamicArray.h:9-16.
 *
 * The r2c / c2r plans are REAL transforms (plain O(n^2)-per-line separable DFTs,
 * FFTW conventions: forward sign -1, unnormalised, in-place padded real layout
 * n0 x n1 x 2(n2/2+1)) so that the reference's spectral curl (Transformer.h:118-295)
 * produces its true vorticity on the small oracle grids.  Single rank only: with
 * NPROCS > 1 a plan is a no-op and says so once.
 */
#pragma once

#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mpi.h>

typedef double fftw_complex[2];

typedef struct mlbm_fft_plan_s {
  int rank;        /* number of dimensions */
  ptrdiff_t n[3];  /* logical (real) sizes, padded with 1s */
  double* real;
  fftw_complex* complex_;
  int backward;
}* fftw_plan;

#define FFTW_ESTIMATE 64u

static inline void* fftw_malloc(size_t bytes) {
  void* pointer = NULL;
  if (posix_memalign(&pointer, 64, bytes ? bytes : 64)) return NULL;
  return pointer;
}
static inline void fftw_free(void* pointer) { free(pointer); }
static inline int fftw_init_threads(void) { return 1; }
static inline void fftw_mpi_init(void) {}
static inline void fftw_plan_with_nthreads(int n) { (void)n; }
static inline void fftw_mpi_cleanup(void) {}

/* local number of COMPLEX elements of the r2c output, slab-decomposed along n[0] */
static inline ptrdiff_t fftw_mpi_local_size(int rank, const ptrdiff_t* n, MPI_Comm comm,
                                            ptrdiff_t* localN0, ptrdiff_t* local0Start) {
  (void)comm;
  int size = 1, me = 0;
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  MPI_Comm_rank(MPI_COMM_WORLD, &me);
  ptrdiff_t total = 1;
  for (int i = 0; i < rank; ++i) total *= (i == rank - 1) ? (n[i] / 2 + 1) : n[i];
  *localN0 = n[0] / size;
  *local0Start = *localN0 * me;
  return total / size;
}

static inline fftw_plan mlbm_fft_make_plan(int rank, const ptrdiff_t* n, double* real,
                                           fftw_complex* complex_, int backward) {
  fftw_plan plan = (fftw_plan)malloc(sizeof(*plan));
  plan->rank = rank;
  for (int i = 0; i < 3; ++i) plan->n[i] = i < rank ? n[i] : 1;
  plan->real = real;
  plan->complex_ = complex_;
  plan->backward = backward;
  return plan;
}

static inline fftw_plan fftw_mpi_plan_dft_r2c(int rank, const ptrdiff_t* n, double* in,
                                              fftw_complex* out, MPI_Comm comm, unsigned flags) {
  (void)comm; (void)flags;
  return mlbm_fft_make_plan(rank, n, in, out, 0);
}

static inline fftw_plan fftw_mpi_plan_dft_c2r(int rank, const ptrdiff_t* n, fftw_complex* in,
                                              double* out, MPI_Comm comm, unsigned flags) {
  (void)comm; (void)flags;
  return mlbm_fft_make_plan(rank, n, out, in, 1);
}

static inline void fftw_destroy_plan(fftw_plan plan) { free(plan); }

/* complex DFT of `count` points with stride `stride` (in complex elements), sign = -1/+1 */
static inline void mlbm_fft_line(fftw_complex* data, ptrdiff_t count, ptrdiff_t stride, int sign,
                                 const double* cosTable, const double* sinTable,
                                 fftw_complex* scratch) {
  for (ptrdiff_t k = 0; k < count; ++k) {
    double re = 0.0, im = 0.0;
    for (ptrdiff_t j = 0; j < count; ++j) {
      ptrdiff_t t = (j * k) % count;
      double c = cosTable[t], s = sign * sinTable[t];
      double xr = data[j * stride][0], xi = data[j * stride][1];
      re += xr * c - xi * s;
      im += xr * s + xi * c;
    }
    scratch[k][0] = re;
    scratch[k][1] = im;
  }
  for (ptrdiff_t k = 0; k < count; ++k) {
    data[k * stride][0] = scratch[k][0];
    data[k * stride][1] = scratch[k][1];
  }
}

static inline void mlbm_fft_tables(ptrdiff_t count, double** cosTable, double** sinTable) {
  *cosTable = (double*)malloc(sizeof(double) * (size_t)count);
  *sinTable = (double*)malloc(sizeof(double) * (size_t)count);
  for (ptrdiff_t t = 0; t < count; ++t) {
    (*cosTable)[t] = cos(2.0 * M_PI * (double)t / (double)count);
    (*sinTable)[t] = sin(2.0 * M_PI * (double)t / (double)count);
  }
}

static inline void fftw_execute(const fftw_plan plan) {
  int size = 1;
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  if (size != 1) {
    static int warned = 0;
    if (!warned) { fprintf(stderr, "mlbm fftw shim: transforms are no-ops with NPROCS > 1\n"); warned = 1; }
    return;
  }
  /* fold leading dimensions so that the transform is always [a][b][last] */
  const int rank = plan->rank;
  const ptrdiff_t last = plan->n[rank - 1];
  const ptrdiff_t half = last / 2 + 1;
  const ptrdiff_t a = rank >= 3 ? plan->n[0] : 1;
  const ptrdiff_t b = rank >= 2 ? plan->n[rank - 2] : 1;
  const ptrdiff_t rows = a * b;
  double *cosLast, *sinLast;
  mlbm_fft_tables(last, &cosLast, &sinLast);
  ptrdiff_t longest = last > a ? last : a;
  if (b > longest) longest = b;
  fftw_complex* scratch = (fftw_complex*)malloc(sizeof(fftw_complex) * (size_t)(longest + 2));
  double* line = (double*)malloc(sizeof(double) * (size_t)(last + 2));

  if (!plan->backward) {
    for (ptrdiff_t r = 0; r < rows; ++r) {
      double* realRow = plan->real + r * 2 * half;
      fftw_complex* complexRow = plan->complex_ + r * half;
      memcpy(line, realRow, sizeof(double) * (size_t)last);
      for (ptrdiff_t k = 0; k < half; ++k) {
        double re = 0.0, im = 0.0;
        for (ptrdiff_t j = 0; j < last; ++j) {
          ptrdiff_t t = (j * k) % last;
          re += line[j] * cosLast[t];
          im -= line[j] * sinLast[t];
        }
        complexRow[k][0] = re;
        complexRow[k][1] = im;
      }
    }
  }

  /* complex transforms along the leading dimensions */
  const int sign = plan->backward ? +1 : -1;
  if (rank >= 2) {
    double *cosB, *sinB;
    mlbm_fft_tables(b, &cosB, &sinB);
    for (ptrdiff_t i = 0; i < a; ++i)
      for (ptrdiff_t k = 0; k < half; ++k)
        mlbm_fft_line(plan->complex_ + i * b * half + k, b, half, sign, cosB, sinB, scratch);
    free(cosB); free(sinB);
  }
  if (rank >= 3) {
    double *cosA, *sinA;
    mlbm_fft_tables(a, &cosA, &sinA);
    for (ptrdiff_t j = 0; j < b; ++j)
      for (ptrdiff_t k = 0; k < half; ++k)
        mlbm_fft_line(plan->complex_ + j * half + k, a, b * half, sign, cosA, sinA, scratch);
    free(cosA); free(sinA);
  }

  if (plan->backward) {
    for (ptrdiff_t r = 0; r < rows; ++r) {
      double* realRow = plan->real + r * 2 * half;
      fftw_complex* complexRow = plan->complex_ + r * half;
      for (ptrdiff_t k = 0; k < half; ++k) { scratch[k][0] = complexRow[k][0]; scratch[k][1] = complexRow[k][1]; }
      for (ptrdiff_t j = 0; j < last; ++j) {
        double value = scratch[0][0];
        for (ptrdiff_t k = 1; k < half; ++k) {
          ptrdiff_t t = (j * k) % last;
          if (2 * k == last) value += scratch[k][0] * cosLast[t];
          else value += 2.0 * (scratch[k][0] * cosLast[t] - scratch[k][1] * sinLast[t]);
        }
        line[j] = value;
      }
      memcpy(realRow, line, sizeof(double) * (size_t)last);
    }
  }
  free(line); free(scratch); free(cosLast); free(sinLast);
}
