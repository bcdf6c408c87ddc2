// step_kernel.cuh -- the fused collide-and-stream PULL kernel (one launch = one lattice step over a
// range of x planes).
//
// Replaces, in one kernel, everything the reference does per node and per step on the device:
//   * the 4 periodic boundary launches (Boundary.h:45-102, Algorithm.h:343-350) -- periodic images are
//     reached by index arithmetic, there are no y/z halo cells at all;
//   * Algorithm::operator() (Algorithm.h:97-126): moments (Moment.h:14-47), force (Force.h:104-292),
//     equilibrium (Equilibrium.h:14-126), entropic alpha (Collision.h:284-375, EntropicStep.h:31-140),
//     forcing source (ForcingScheme.h:41-198), collide-and-stream (Collision.h:134-151, 243-258);
//   * Algorithm::storeFields (Algorithm.h:150-194) and the per-node part of the scalar analyses
//     (Analysis.h:53-61) on stored steps, as warp-shuffle + block reductions.
//
// Memory behaviour (HBM-bound design): every population is read exactly once (pulled from its upstream
// neighbour, coalesced along the unit-stride axis r) and written exactly once (perfectly aligned), so the
// algorithmic traffic is 2*Q*sizeof(StoreT) bytes per node (+2*sizeof(StoreT) for the ELBM alpha field).
// All Q loads of a node are issued before the first use, which keeps >= Q independent requests in flight
// per thread.  Arithmetic is FP64 in registers whatever the storage type.
#pragma once

#include <cuda_runtime.h>

#include <type_traits>

#include "lattice.cuh"

namespace mlbm {

enum CollisionKind { kBGK = 0, kELBM = 1 };
enum EquilibriumKind { kTruncationMa3 = 0, kExact = 1 };
// ShanChen has a zero collision source (ForcingScheme.h:141-151) and therefore shares the kernel of
// "None"; the two only differ in the stored hydrodynamic velocity (hydroShift below).
enum SchemeKind { kSchemeNone = 0, kSchemeGuo = 1, kSchemeEDM = 2 };

constexpr int kStepBlock = 128;
constexpr int kObservableSlots = 3;  // energy, mass, max |u|^2

struct StepParams {
  const void* prev;        // SoA populations read by this step  [Q][LX+2][NM][NR]
  void* next;              // SoA populations written by this step
  void* alpha;             // [LX][NM][NR], ELBM warm start, read and written every step (Algorithm.h:103-106)
  void* density;           // stored fields (only touched when isStored)
  void* velocity;          // [D] components, fieldStride apart
  void* force;             // [D] components, fieldStride apart
  double* partials;        // [LX * NM * ceil(NR / kStepBlock)][kObservableSlots] block partial sums (only when isStored)
  const double* forceTable[3];  // per force component: amplitude * profile along forceAxis (host libm values)
  int forceAxis[3];        // 0 = x, 1 = m, 2 = r, -1 = component is identically zero
  long long stride;        // elements between populations
  long long plane;         // elements between x planes (= NM * NR)
  long long fieldStride;   // elements between field components
  int LX, NM, NR;          // local interior extents along x, m, r
  int x0;                  // first local x plane of this launch (blockIdx.z counts from it)
  int wrapX;               // 1: single rank, x is periodic inside the slab; 0: halo planes hold the neighbours' data
  int isStored;            // Algorithm::isStored (Routine.h:122-124): bit 0 = store fields, bit 1 = reduce observables
  int hydroShift;          // 1: stored velocity = u + F/(2 rho) (ForcingScheme.h:26-33); 0: u (scheme None, :50-57)
  int hasForce;            // 0: force is identically zero
  double beta;             // 1 / (2 tau)                      (Collision.h:122)
  double guoFactor;        // (1 - 1/(2 tau)) * inv_cs2        (ForcingScheme.h:115)
};

template <typename StoreT> __device__ __forceinline__ double loadPopulation(const StoreT* p) {
  return (double)__ldg(p);
}
template <typename StoreT> __device__ __forceinline__ void storePopulation(StoreT* p, double v) {
  __stcs(p, (StoreT)v);
}

// ------------------------------------------------------------------------------------------------
// Equilibrium, evaluated for all Q populations from q-independent coefficients.
//   TruncationMa3 (Equilibrium.h:17-34): the reference's 9-term polynomial regrouped by powers of c.u
//     P = A0 + cu (A1 + cu (A2 + cu (A3 + cu A4))),  s = inv_cs2 = 3
//     A0 = 1 - s/2 u2 + s^2/8 u2^2, A1 = s - s^2/2 u2, A2 = s^2/2 - s^3/4 u2, A3 = s^3/6, A4 = s^4/24
//   Exact (Equilibrium.h:60-81, 106-126): product form, three factors per dimension precomputed.
// ------------------------------------------------------------------------------------------------
template <class L, int EQ> struct EquilibriumCoefficients;

template <class L> struct EquilibriumCoefficients<L, kTruncationMa3> {
  double a0, a1, a2;
  double u[3];
  __device__ __forceinline__ void set(const double* velocity, double u2) {
    a0 = 1.0 - 1.5 * u2 + 1.125 * u2 * u2;
    a1 = 3.0 - 4.5 * u2;
    a2 = 4.5 - 6.75 * u2;
#pragma unroll
    for (int d = 0; d < 3; ++d) u[d] = d < L::D ? velocity[d] : 0.0;
  }
  // returns feq / (rho * w_q)
  template <int q> __device__ __forceinline__ double shape() const {
    double cu = 0.0;
#pragma unroll
    for (int d = 0; d < L::D; ++d) {
      if (L::c(q, d) == 1) cu += u[d];
      if (L::c(q, d) == -1) cu -= u[d];
    }
    if (L::norm2(q) == 0) return a0;
    return a0 + cu * (a1 + cu * (a2 + cu * (4.5 + cu * 3.375)));
  }
};

template <class L> struct EquilibriumCoefficients<L, kExact> {
  double factor[3][3];  // [d][c+1]: (2 - sqrt(1+3u^2)) * ((2u + sqrt(1+3u^2)) / (1-u))^c
  __device__ __forceinline__ void set(const double* velocity, double) {
#pragma unroll
    for (int d = 0; d < L::D; ++d) {
      const double ud = velocity[d];
      const double root = sqrt(1.0 + 3.0 * ud * ud);
      const double a = 2.0 - root;
      const double b = (2 * ud + root) / (1.0 - ud);
      factor[d][0] = a * (1.0 / b);
      factor[d][1] = a;
      factor[d][2] = a * b;
    }
  }
  template <int q> __device__ __forceinline__ double shape() const {
    double r = factor[0][L::c(q, 0) + 1];
#pragma unroll
    for (int d = 1; d < L::D; ++d) r *= factor[d][L::c(q, d) + 1];
    return r;
  }
};

template <int I, int N, class F> __device__ __forceinline__ void staticFor(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    staticFor<I + 1, N>(f);
  }
}

// ------------------------------------------------------------------------------------------------
// fastLog: natural logarithm of a positive normal double with an absolute error of a few 1e-17 * (1 + |ln v|),
// about 12 FP64 instructions instead of the ~30 FP64 + ~25 integer instructions of the CUDA math library's log().
// The entropic solve evaluates (1 + iterations) * Q logarithms per node, which makes it FP64-pipe bound (SURVEY.md
// section 7); this is what moves it back towards the HBM roofline.
//   v = 2^k z, z in [0.6875, 1.375): integer arithmetic on the bit pattern; the top 7 bits below the exponent pick
//   one of 128 sub-intervals with centre c; invc = double(1/c), logc = double(-ln invc) (log_table.inc, generated with
//   100-digit arithmetic by gen_log_table.py); r = fma(z, invc, -1) is exact to rounding and |r| < 2^-8, so
//   ln v = k ln2 + logc + (r - r^2/2 + ... + r^7/7) with a truncation error below 1e-20.
// Zero, negative, subnormal, infinite and NaN arguments take the library path so that the Newton iteration sees the
// same NaNs the reference's std::log produces for a mirror state that left the positive cone (EntropicStep.h:31-62).
// ------------------------------------------------------------------------------------------------
static __device__ const double2 kLogTable[128] = {
#include "log_table.inc"
};

static __device__ __noinline__ double libraryLog(double v) { return log(v); }

__device__ __forceinline__ double fastLog(double v, const double2* __restrict__ table) {
  if (!(v >= 2.2250738585072014e-308 && v <= 1.7976931348623157e308)) return libraryLog(v);
  const long long ix = __double_as_longlong(v);
  const long long tmp = ix - 0x3FE6000000000000LL;
  const double kd = (double)(int)(tmp >> 52);
  const int i = (int)(tmp >> 45) & 127;
  const double z = __longlong_as_double(ix - (tmp & (long long)0xFFF0000000000000ULL));
  const double2 entry = table[i];
  const double r = fma(z, entry.x, -1.0);
  double p = fma(r, 1.0 / 7.0, -1.0 / 6.0);
  p = fma(r, p, 0.2);
  p = fma(r, p, -0.25);
  p = fma(r, p, 1.0 / 3.0);
  p = fma(r, p, -0.5);
  const double hi = fma(kd, 0x1.62e42fefa3800p-1, entry.y);
  return hi + fma(r * r, p, fma(kd, 0x1.ef35793c76730p-45, r));
}

// ------------------------------------------------------------------------------------------------
// Entropic alpha: Collision<ELBM>::calculateAlpha (Collision.h:351-375).
// ------------------------------------------------------------------------------------------------
template <class L>
__device__ __forceinline__ double entropicAlpha(const double (&f)[L::Q], const double (&fNeq)[L::Q], double alphaGuess,
                                                const double2* __restrict__ logTable) {
  // isDeviationSmall (Collision.h:284-303): no |fNeq_q| / f_q above 1e-3
  bool small = true;
#pragma unroll
  for (int q = 0; q < L::Q; ++q) {
    const double a = fabs(fNeq[q]);
    const bool large = f[q] > 0.0 ? (a > 1.0e-3 * f[q]) : (f[q] == 0.0 ? a > 0.0 : false);
    small = small && !large;
  }
  if (small) return 2.0;

  // calculateAlphaMax (Collision.h:305-326): min(2.5, min over fNeq_q > 0 of |f_q| / fNeq_q), tracked as a fraction
  double num = 2.5, den = 1.0;
#pragma unroll
  for (int q = 0; q < L::Q; ++q) {
    if (fNeq[q] > 0.0) {
      const double a = fabs(f[q]);
      if (a * den < num * fNeq[q]) { num = a; den = fNeq[q]; }
    }
  }
  const double alphaMax = num / den;
  if (alphaMax < 2.0) return 0.95 * alphaMax;

  // solveAlpha (Collision.h:328-349) -> NewtonRaphsonSolver (EntropicStep.h:111-140) on
  //   F(a)  = sum f ln(f/w) - (f - a fNeq) ln((f - a fNeq)/w)      (EntropicStep.h:31-45)
  //   F'(a) = sum fNeq (1 + ln((f - a fNeq)/w))                     (EntropicStep.h:47-62)
  // The a-independent sum is hoisted and ln((f - a fNeq)/w) is shared between F and F'.
  double hoisted = 0.0;
#pragma unroll
  for (int q = 0; q < L::Q; ++q) hoisted = fma(f[q], fastLog(f[q] * (1.0 / L::w(q)), logTable), hoisted);

  double x = alphaGuess, step = 0.0;
  bool converged = false;
  for (int iteration = 1; iteration <= 50; ++iteration) {
    x = x - step;
    double sum = 0.0, derivative = 0.0;
#pragma unroll
    for (int q = 0; q < L::Q; ++q) {
      const double g = fma(-x, fNeq[q], f[q]);
      const double lg = fastLog(g * (1.0 / L::w(q)), logTable);
      sum = fma(g, lg, sum);
      derivative = fma(fNeq[q], 1.0 + lg, derivative);
    }
    step = (hoisted - sum) / derivative;
    if (fabs(step) <= 1e-8) { converged = (x > 1.0 && x < alphaMax); break; }
  }
  return converged ? x : 2.0;
}

// ------------------------------------------------------------------------------------------------
// The fused step.
// grid = (ceil(NR / kStepBlock), NM, number of x planes), block = kStepBlock threads along r.
// ------------------------------------------------------------------------------------------------
template <class L, int COLLISION, int EQ, int SCHEME, typename StoreT>
__global__ void __launch_bounds__(kStepBlock)
fusedStepKernel(const __grid_constant__ StepParams p) {
  constexpr int Q = L::Q;
  constexpr int D = L::D;

  const int r = blockIdx.x * kStepBlock + threadIdx.x;
  const int m = blockIdx.y;
  const int x = p.x0 + blockIdx.z;
  const bool active = r < p.NR;

  __shared__ double2 logTable[COLLISION == kELBM ? 128 : 1];
  if (COLLISION == kELBM) {
    static_assert(kStepBlock == 128, "one table entry per thread");
    logTable[threadIdx.x] = kLogTable[threadIdx.x];
    __syncthreads();
  }

  double rho = 0.0, energy = 0.0, speed2 = 0.0;

  if (active) {
    // upstream coordinates: pull from (x - cx, m - cm, r - cr) of the periodic image
    const int xh = x + 1;  // halo plane 0 precedes the interior
    int xPrev = xh - 1, xNext = xh + 1;
    if (p.wrapX) {
      if (xPrev == 0) xPrev = p.LX;
      if (xNext == p.LX + 1) xNext = 1;
    }
    const int mPrev = m == 0 ? p.NM - 1 : m - 1;
    const int mNext = m == p.NM - 1 ? 0 : m + 1;
    const int rPrev = r == 0 ? p.NR - 1 : r - 1;
    const int rNext = r == p.NR - 1 ? 0 : r + 1;

    const StoreT* __restrict__ prev = static_cast<const StoreT*>(p.prev);
    StoreT* __restrict__ next = static_cast<StoreT*>(p.next);

    double f[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int xs = L::cx(q) == 1 ? xPrev : (L::cx(q) == -1 ? xNext : xh);
      const int ms = L::cm(q) == 1 ? mPrev : (L::cm(q) == -1 ? mNext : m);
      const int rs = L::cr(q) == 1 ? rPrev : (L::cr(q) == -1 ? rNext : r);
      f[q] = loadPopulation(prev + q * p.stride + xs * p.plane + (long long)ms * p.NR + rs);
    }

    // Moment::calculateDensity / calculateVelocity (Moment.h:14-47)
    rho = f[0];
#pragma unroll
    for (int q = 1; q < Q; ++q) rho += f[q];
    double u[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int q = 1; q < Q; ++q) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        if (L::c(q, d) == 1) u[d] += f[q];
        if (L::c(q, d) == -1) u[d] -= f[q];
      }
    }
    const double invRho = 1.0 / rho;
    double u2 = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      u[d] *= invRho;
      u2 += u[d] * u[d];
    }

    // Force::setForce at local interior coordinates (Collision.h:81-88); profiles precomputed on the host
    double F[3] = {0.0, 0.0, 0.0};
    if (p.hasForce) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const int axis = p.forceAxis[d];
        if (axis >= 0) F[d] = __ldg(p.forceTable[d] + (axis == 0 ? x : (axis == 1 ? m : r)));
      }
    }

    EquilibriumCoefficients<L, EQ> eq;
    eq.set(u, u2);

    const long long node = (long long)x * p.plane + (long long)m * p.NR + r;  // field / alpha index
    const long long out = (long long)xh * p.plane + (long long)m * p.NR + r;

    // collision source helpers
    double uF = 0.0;
    EquilibriumCoefficients<L, EQ> eqShifted;  // EDM: feq at u + F / rho (ForcingScheme.h:184-197)
    if (SCHEME == kSchemeGuo) {
#pragma unroll
      for (int d = 0; d < D; ++d) uF += u[d] * F[d];
    }
    if (SCHEME == kSchemeEDM) {
      double v[3] = {0.0, 0.0, 0.0};
      double v2 = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        v[d] = u[d] + F[d] * invRho;
        v2 += v[d] * v[d];
      }
      eqShifted.set(v, v2);
    }

    double alpha = 2.0;
    if (COLLISION == kELBM) {
      // Collision<ELBM>::calculateRelaxationTime (Collision.h:227-241)
      double fNeq[Q];
      staticFor<0, Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        fNeq[q] = f[q] - rho * L::w(q) * eq.template shape<q>();
      });
      StoreT* alphaField = static_cast<StoreT*>(p.alpha);
      alpha = entropicAlpha<L>(f, fNeq, (double)alphaField[node], logTable);
      alphaField[node] = (StoreT)alpha;
      const double omega = alpha * p.beta;  // 1 / tau_eff (Collision.h:240)
      // Collision<ELBM>::collideAndStream (Collision.h:243-258)
      staticFor<0, Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        double value = f[q] - omega * fNeq[q];
        if (SCHEME == kSchemeGuo) {
          double cF = 0.0, cu = 0.0;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            if (L::c(q, d) == 1) { cF += F[d]; cu += u[d]; }
            if (L::c(q, d) == -1) { cF -= F[d]; cu -= u[d]; }
          }
          value += p.guoFactor * L::w(q) * (cF - uF + 3.0 * cu * cF);
        }
        if (SCHEME == kSchemeEDM) {
          value += rho * L::w(q) * eqShifted.template shape<q>() - (f[q] - fNeq[q]);
        }
        storePopulation(next + q * p.stride + out, value);
      });
    } else {
      // Collision<BGK>::collideAndStream (Collision.h:134-151)
      const double keep = 1.0 - 2.0 * p.beta;
      const double relax = 2.0 * p.beta;
      staticFor<0, Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        const double feq = rho * L::w(q) * eq.template shape<q>();
        double value = keep * f[q] + relax * feq;
        if (SCHEME == kSchemeGuo) {
          double cF = 0.0, cu = 0.0;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            if (L::c(q, d) == 1) { cF += F[d]; cu += u[d]; }
            if (L::c(q, d) == -1) { cF -= F[d]; cu -= u[d]; }
          }
          value += p.guoFactor * L::w(q) * (cF - uF + 3.0 * cu * cF);
        }
        if (SCHEME == kSchemeEDM) {
          value += rho * L::w(q) * eqShifted.template shape<q>() - feq;
        }
        storePopulation(next + q * p.stride + out, value);
      });
    }

    if (p.isStored) {
      // Algorithm::storeFields (Algorithm.h:150-194); BGK's alpha field is the constant 2 (Collision.h:121)
      const bool fields = (p.isStored & 1) != 0;
      if (fields) static_cast<StoreT*>(p.density)[node] = (StoreT)rho;
      const double half = p.hydroShift ? 0.5 * invRho : 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double v = u[d] + half * F[d];
        if (fields) {
          static_cast<StoreT*>(p.velocity)[d * p.fieldStride + node] = (StoreT)v;
          static_cast<StoreT*>(p.force)[d * p.fieldStride + node] = (StoreT)F[d];
        }
        energy += 0.5 * rho * v * v;  // TotalEnergy (Analysis.h:53-61)
        speed2 += v * v;
      }
    }
  }

  if (p.isStored) {
    // block reduction: warp shuffles, then one value per warp through shared memory
    double mass = active ? rho : 0.0;
#pragma unroll
    for (int offset = 16; offset > 0; offset >>= 1) {
      energy += __shfl_xor_sync(0xffffffffu, energy, offset);
      mass += __shfl_xor_sync(0xffffffffu, mass, offset);
      speed2 = fmax(speed2, __shfl_xor_sync(0xffffffffu, speed2, offset));
    }
    __shared__ double scratch[kObservableSlots][kStepBlock / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
      scratch[0][warp] = energy;
      scratch[1][warp] = mass;
      scratch[2][warp] = speed2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double e = 0.0, ms = 0.0, s2 = 0.0;
#pragma unroll
      for (int i = 0; i < kStepBlock / 32; ++i) {
        e += scratch[0][i];
        ms += scratch[1][i];
        s2 = fmax(s2, scratch[2][i]);
      }
      const long long block = ((long long)x * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      p.partials[block * kObservableSlots + 0] = e;
      p.partials[block * kObservableSlots + 1] = ms;
      p.partials[block * kObservableSlots + 2] = s2;
    }
  }
}

using StepKernel = void (*)(const StepParams);

// Returns the specialised kernel for a configuration or nullptr when the reference has no such
// combination (e.g. the exact equilibrium exists only for D2Q9 and D3Q27, Equilibrium.h:36-126).
// Defined across the instantiate_*.cu translation units.
StepKernel lookupStepKernel(int lattice, int collision, int equilibrium, int scheme, int dtype);

}  // namespace mlbm
