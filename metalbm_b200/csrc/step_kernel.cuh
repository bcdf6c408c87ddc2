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
  int x0;                  // first local x plane of this launch
  int planeStep;           // blockIdx.z-th plane of the launch is x0 + blockIdx.z * planeStep (1: a contiguous range;
                           // LX - 1 with two planes: the two boundary planes of the slab in one launch)
  void* peerLow;           // direct peer halos: the LEFT neighbour's `next` buffer, mapped over NVLink;
                           // plane x = 0 stores its c_x < 0 populations into that buffer's halo plane LX + 1 as well
  void* peerHigh;          // the RIGHT neighbour's `next` buffer; plane x = LX - 1 stores its c_x > 0 populations into
                           // that buffer's halo plane 0 (what Communication.h:134-180 sends); nullptr = no peer stores
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
// fastLogCore: natural logarithm of doubles in [0.25, 4) with an absolute error below 2.5e-16, 9 FP64 + 4 integer
// instructions instead of the ~30 FP64 + ~25 integer instructions of the CUDA math library's log().
// The entropic solve evaluates (1 + iterations) * Q logarithms per node, which makes it FP64-issue bound (SURVEY.md
// section 7); this is what moves it back towards the HBM roofline.  The arguments are f_q / w_q and
// (f_q - alpha fNeq_q) / w_q, i.e. the local density times 1 + O(Mach) + O(non-equilibrium): [0.25, 4) covers every
// state a lattice-Boltzmann run can sensibly be in, and anything else takes the library path (entropicNewtonLibrary).
//   The top bits of the double (exponent and 7 mantissa bits, minus those of 0.25) index one of 512 sub-intervals with
//   centre c; invc = double(1/c), logc = double(-ln invc) (log_table.inc, generated with 100-digit arithmetic by
//   gen_log_table.py); r = fma(v, invc, -1) is exact to rounding and |r| < 2^-8, so
//   ln v = logc + (r - r^2/2 + ... + r^7/7) with a truncation error below 1e-20.
// N independent arguments advance in lock step: every Horner step is issued for all N before the next one, which
// gives the FP64 pipe N independent dependency chains per warp.
// `range` accumulates the maximum table index as an unsigned number: it stays below 512 exactly when every argument
// was inside [0.25, 4) (smaller, negative, infinite and NaN arguments all map to indices >= 512).
// ------------------------------------------------------------------------------------------------
constexpr int kLogTableEntries = 512;
static __device__ const double2 kLogTable[kLogTableEntries] = {
#include "log_table.inc"
};

template <int N>
__device__ __forceinline__ void fastLogCore(const double (&v)[N], double (&out)[N], const double2* __restrict__ table,
                                            unsigned& range) {
  double2 entry[N];
  double r[N], p[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const unsigned index = (unsigned)((__double2hiint(v[j]) - 0x3FD00000) >> 13);
    range = max(range, index);
    entry[j] = table[index & (kLogTableEntries - 1)];
  }
#pragma unroll
  for (int j = 0; j < N; ++j) r[j] = fma(v[j], entry[j].x, -1.0);
#pragma unroll
  for (int j = 0; j < N; ++j) p[j] = fma(r[j], 1.0 / 7.0, -1.0 / 6.0);
#pragma unroll
  for (int j = 0; j < N; ++j) p[j] = fma(r[j], p[j], 0.2);
#pragma unroll
  for (int j = 0; j < N; ++j) p[j] = fma(r[j], p[j], -0.25);
#pragma unroll
  for (int j = 0; j < N; ++j) p[j] = fma(r[j], p[j], 1.0 / 3.0);
#pragma unroll
  for (int j = 0; j < N; ++j) p[j] = fma(r[j], p[j], -0.5);
#pragma unroll
  for (int j = 0; j < N; ++j) out[j] = entry[j].y + fma(r[j] * r[j], p[j], r[j]);
}

// general entry point (self-test, tools): library logarithm outside [0.25, 4)
__device__ __forceinline__ double fastLog(double v, const double2* __restrict__ table) {
  const double in[1] = {v};
  double out[1];
  unsigned range = 0;
  fastLogCore<1>(in, out, table, range);
  return range < kLogTableEntries ? out[0] : log(v);
}

// ------------------------------------------------------------------------------------------------
// Entropic alpha: Collision<ELBM>::calculateAlpha (Collision.h:351-375).
//
// Work layout.  The populations f_q and their non-equilibrium parts of one node live in SHARED memory while alpha is
// solved for ([q][thread], conflict-free), not in registers: the Newton loops over q are then rolled (a few dozen
// instructions, three-way unrolled for instruction-level parallelism) instead of Q copies of the logarithm per
// evaluation, which had overflowed the instruction cache (ncu: 'no_instructions' was the second stall reason) and
// kept 108 registers alive through the solve.
// ------------------------------------------------------------------------------------------------
struct EntropicScratch {
  const double* f;       // f[q * kStepBlock]     (this thread's column)
  const double* fNeq;    // fNeq[q * kStepBlock]
  const double* invW;    // 1 / w_q, one copy per block
  const double2* table;  // fastLog table (shared or global memory)
};

// The same solve with the CUDA math library's logarithm, for the rare node whose arguments fall outside fastLogCore's
// table.  Kept out of line and rolled: it is cold code.
template <int Q>
__device__ __noinline__ double entropicNewtonLibrary(const EntropicScratch& s, double alphaGuess, double alphaMax) {
  double hoisted = 0.0;
#pragma unroll 1
  for (int q = 0; q < Q; ++q) {
    const double fq = s.f[q * kStepBlock];
    hoisted = fma(fq, log(fq * s.invW[q]), hoisted);
  }
  double x = alphaGuess, step = 0.0;
  for (int iteration = 1; iteration <= 50; ++iteration) {
    x = x - step;
    double sum = 0.0, derivative = 0.0;
#pragma unroll 1
    for (int q = 0; q < Q; ++q) {
      const double nq = s.fNeq[q * kStepBlock];
      const double g = fma(-x, nq, s.f[q * kStepBlock]);
      const double lg = log(g * s.invW[q]);
      sum = fma(g, lg, sum);
      derivative = fma(nq, 1.0 + lg, derivative);
    }
    step = (hoisted - sum) / derivative;
    if (fabs(step) <= 1e-8) return (x > 1.0 && x < alphaMax) ? x : 2.0;
  }
  return 2.0;
}

// The q loops are rolled over groups of three populations (the three logarithms of a group advance in lock step,
// see fastLogCore); Q = 9, 15, 27 are multiples of three, the others end with a partial group whose unused slots
// are fed the harmless argument 1 and contribute exact zeros.
template <int Q>
__device__ __forceinline__ double entropicNewton(const EntropicScratch& s, double alphaGuess, double alphaMax) {
  // solveAlpha (Collision.h:328-349) -> NewtonRaphsonSolver (EntropicStep.h:111-140) on
  //   F(a)  = sum f ln(f/w) - (f - a fNeq) ln((f - a fNeq)/w)      (EntropicStep.h:31-45)
  //   F'(a) = sum fNeq (1 + ln((f - a fNeq)/w))                     (EntropicStep.h:47-62)
  // The a-independent sum is hoisted and ln((f - a fNeq)/w) is shared between F and F'.  Each slot of a group keeps
  // its own partial sums.
  // As soon as a logarithm argument leaves the table's range the node is handed to entropicNewtonLibrary, which
  // restarts the solve with the library logarithm (and with it the reference's NaN behaviour for mirror states that
  // leave the positive cone: the iteration produces NaNs, gives up after 50 iterations and alpha falls back to 2,
  // EntropicStep.h:126-138, Collision.h:344-346).
  constexpr int G = 3;
  constexpr int groups = (Q + G - 1) / G;
  unsigned range = 0;
  double h[G] = {0.0, 0.0, 0.0};
#pragma unroll 1
  for (int group = 0; group < groups; ++group) {
    double fq[G], v[G], lg[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int q = group * G + j;
      const bool live = Q % G == 0 || q < Q;
      fq[j] = live ? s.f[q * kStepBlock] : 0.0;
      v[j] = live ? fq[j] * s.invW[q] : 1.0;
    }
    fastLogCore<G>(v, lg, s.table, range);
#pragma unroll
    for (int j = 0; j < G; ++j) h[j] = fma(fq[j], lg[j], h[j]);
  }
  if (range >= kLogTableEntries) return entropicNewtonLibrary<Q>(s, alphaGuess, alphaMax);
  const double hoisted = (h[0] + h[1]) + h[2];

  double x = alphaGuess, step = 0.0;
  bool converged = false;
  for (int iteration = 1; iteration <= 50; ++iteration) {
    x = x - step;
    double sum[G] = {0.0, 0.0, 0.0}, derivative[G] = {0.0, 0.0, 0.0};
#pragma unroll 1
    for (int group = 0; group < groups; ++group) {
      double g[G], nq[G], v[G], lg[G];
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const int q = group * G + j;
        const bool live = Q % G == 0 || q < Q;
        nq[j] = live ? s.fNeq[q * kStepBlock] : 0.0;
        g[j] = live ? fma(-x, nq[j], s.f[q * kStepBlock]) : 0.0;
        v[j] = live ? g[j] * s.invW[q] : 1.0;
      }
      fastLogCore<G>(v, lg, s.table, range);
#pragma unroll
      for (int j = 0; j < G; ++j) {
        sum[j] = fma(g[j], lg[j], sum[j]);
        derivative[j] = fma(nq[j], 1.0 + lg[j], derivative[j]);
      }
    }
    if (range >= kLogTableEntries) return entropicNewtonLibrary<Q>(s, alphaGuess, alphaMax);
    step = (hoisted - ((sum[0] + sum[1]) + sum[2])) / ((derivative[0] + derivative[1]) + derivative[2]);
    if (fabs(step) <= 1e-8) { converged = (x > 1.0 && x < alphaMax); break; }
  }
  return converged ? x : 2.0;
}

// bytes of dynamic shared memory the entropic kernels need (see fusedStepKernel)
constexpr int kLogTableBytes = kLogTableEntries * 16;
constexpr int entropicSharedBytes(int Q, bool tableInShared) {
  return 2 * Q * kStepBlock * 8 + ((Q * 8 + 15) / 16) * 16 + (tableInShared ? kLogTableBytes : 0);
}
// the fastLog table is staged in shared memory whenever four blocks per SM still fit (228 KB, 1 KB reserved per block)
constexpr bool logTableInShared(int Q) { return 4 * (entropicSharedBytes(Q, true) + 1024) <= 233472; }

// ------------------------------------------------------------------------------------------------
// The fused step.
// grid = (ceil(NR / kStepBlock), NM, number of x planes), block = kStepBlock threads along r.
// ------------------------------------------------------------------------------------------------
template <class L, int COLLISION, int EQ, int SCHEME, typename StoreT>
__global__ void __launch_bounds__(kStepBlock, COLLISION == kELBM ? 4 : 1)
fusedStepKernel(const __grid_constant__ StepParams p) {
  constexpr int Q = L::Q;
  constexpr int D = L::D;

  const int r = blockIdx.x * kStepBlock + threadIdx.x;
  const int m = blockIdx.y;
  const int x = p.x0 + (int)blockIdx.z * p.planeStep;
  const bool active = r < p.NR;

  // entropic kernels: dynamic shared memory = f[Q][block] | fNeq[Q][block] | 1/w[Q] | (fastLog table)
  extern __shared__ __align__(16) unsigned char dynamicShared[];
  EntropicScratch scratchPointers = {nullptr, nullptr, nullptr, kLogTable};
  double* sharedF = nullptr;
  double* sharedFNeq = nullptr;
  if (COLLISION == kELBM) {
    static_assert(kLogTableEntries % kStepBlock == 0, "whole table entries per thread");
    sharedF = reinterpret_cast<double*>(dynamicShared) + threadIdx.x;
    sharedFNeq = sharedF + Q * kStepBlock;
    double* invW = reinterpret_cast<double*>(dynamicShared) + 2 * Q * kStepBlock;
    staticFor<0, Q>([&](auto qc) {
      constexpr int q = decltype(qc)::value;
      if (threadIdx.x == q) invW[q] = 1.0 / L::w(q);
    });
    scratchPointers.f = sharedF;
    scratchPointers.fNeq = sharedFNeq;
    scratchPointers.invW = invW;
    if (logTableInShared(Q)) {
      double2* table = reinterpret_cast<double2*>(dynamicShared + 2 * Q * kStepBlock * 8 + ((Q * 8 + 15) / 16) * 16);
#pragma unroll
      for (int i = 0; i < kLogTableEntries / kStepBlock; ++i) table[i * kStepBlock + threadIdx.x] = kLogTable[i * kStepBlock + threadIdx.x];
      scratchPointers.table = table;
    }
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
    // halo planes of the neighbours this node's outgoing populations belong to (block-uniform conditions)
    StoreT* const remoteHigh = (p.peerHigh && x == p.LX - 1) ? static_cast<StoreT*>(p.peerHigh) + ((long long)m * p.NR + r) : nullptr;
    StoreT* const remoteLow = (p.peerLow && x == 0) ? static_cast<StoreT*>(p.peerLow) + ((long long)(p.LX + 1) * p.plane + (long long)m * p.NR + r) : nullptr;
    auto storeOutgoing = [&](auto qc, double value) {
      constexpr int q = decltype(qc)::value;
      storePopulation(next + q * p.stride + out, value);
      if (L::cx(q) == 1 && remoteHigh) remoteHigh[q * p.stride] = (StoreT)value;
      if (L::cx(q) == -1 && remoteLow) remoteLow[q * p.stride] = (StoreT)value;
    };

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
      // Collision<ELBM>::calculateRelaxationTime (Collision.h:227-241): fNeq, then alpha.  While fNeq is formed the
      // two cheap screens of calculateAlpha run on the register values:
      //   isDeviationSmall (Collision.h:284-303): no |fNeq_q| / f_q above 1e-3
      //   calculateAlphaMax (Collision.h:305-326): min(2.5, min over fNeq_q > 0 of |f_q| / fNeq_q), tracked as a fraction
      bool small = true;
      double num = 2.5, den = 1.0;
      staticFor<0, Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        const double fq = f[q];
        const double nq = fq - rho * L::w(q) * eq.template shape<q>();
        sharedF[q * kStepBlock] = fq;
        sharedFNeq[q * kStepBlock] = nq;
        const double a = fabs(nq);
        const bool large = fq > 0.0 ? (a > 1.0e-3 * fq) : (fq == 0.0 ? a > 0.0 : false);
        small = small && !large;
        if (nq > 0.0) {
          const double af = fabs(fq);
          if (af * den < num * nq) { num = af; den = nq; }
        }
      });
      StoreT* alphaField = static_cast<StoreT*>(p.alpha);
      if (!small) {
        // Collision<ELBM>::calculateAlpha (Collision.h:351-375)
        const double alphaMax = num / den;
        alpha = alphaMax < 2.0 ? 0.95 * alphaMax : entropicNewton<Q>(scratchPointers, (double)alphaField[node], alphaMax);
      }
      alphaField[node] = (StoreT)alpha;
      const double omega = alpha * p.beta;  // 1 / tau_eff (Collision.h:240)
      // Collision<ELBM>::collideAndStream (Collision.h:243-258)
      staticFor<0, Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        const double fq = sharedF[q * kStepBlock], nq = sharedFNeq[q * kStepBlock];
        double value = fq - omega * nq;
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
          value += rho * L::w(q) * eqShifted.template shape<q>() - (fq - nq);
        }
        storeOutgoing(qc, value);
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
        storeOutgoing(qc, value);
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
