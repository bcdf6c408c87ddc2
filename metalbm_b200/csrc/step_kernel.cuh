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

// kELBMForcing = Collision<ForcedNR_ELBM_Forcing> (Collision.h:727-857): the entropic solve on the FORCED populations
enum CollisionKind { kBGK = 0, kELBM = 1, kELBMForcing = 2 };
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
  unsigned long long* newtonCounters;  // entropic kernels: [0] += nodes that took the Newton solve, [1] += evaluations of (F, F') they needed
  const double* forceTable[3];  // per force component: amplitude * profile along forceAxis (host libm values)
  int forceAxis[3];        // 0 = x, 1 = m, 2 = r, -1 = component is identically zero
  long long stride;        // elements between populations
  long long plane;         // elements between x planes (= NM * NR)
  long long fieldStride;   // elements between field components
  int LX, NM, NR;          // local interior extents along x, m, r
  int x0;                  // first local x plane of this launch
  int planeStep;           // blockIdx.z-th plane of the launch is x0 + blockIdx.z * planeStep (1: a contiguous range;
                           // LX - 1 with two planes: the two boundary planes of the slab in one launch)
  int planeCount;          // number of planes of the launch
  int planesPerBlock;      // entropic kernels: consecutive planes walked by one block (gridDim.z = ceil(planeCount / planesPerBlock))
  void* peerLow;           // direct peer halos: the LEFT neighbour's `next` buffer, mapped over NVLink;
                           // plane x = 0 stores its c_x < 0 populations into that buffer's halo plane LX + 1 as well
  void* peerHigh;          // the RIGHT neighbour's `next` buffer; plane x = LX - 1 stores its c_x > 0 populations into
                           // that buffer's halo plane 0 (what Communication.h:134-180 sends); nullptr = no peer stores
  int wrapX;               // 1: single rank, x is periodic inside the slab; 0: halo planes hold the neighbours' data
  int isStored;            // Algorithm::isStored (Routine.h:122-124): bit 0 = store fields, bit 1 = reduce observables
  int hydroShift;          // 1: stored velocity = u + F/(2 rho) (ForcingScheme.h:26-33); 0: u (scheme None, :50-57)
  int hasForce;            // 0: force is identically zero; 1: per-axis profiles (forceTable); 2: read from the force FIELD
                           // (`force`, the generic array read of Force.h:39-48 that the spectral forces of Force.h:296-623 use)
  double beta;             // 1 / (2 tau)                      (Collision.h:122)
  double guoFactor;        // (1 - 1/(2 tau)) * inv_cs2        (ForcingScheme.h:115)
};

template <typename CT = double, typename StoreT> __device__ __forceinline__ CT loadPopulation(const StoreT* p) {
  return (CT)__ldg(p);
}
// L2-only load: the entropic kernels keep their logarithm table (and little else) in what the shared-memory carve-out
// leaves of L1; the once-read population stream must not evict it
template <typename CT = double, typename StoreT> __device__ __forceinline__ CT loadPopulationStreaming(const StoreT* p) {
  return (CT)__ldcg(p);
}
template <typename StoreT, typename CT> __device__ __forceinline__ void storePopulation(StoreT* p, CT v) {
  __stcs(p, (StoreT)v);
}

// ------------------------------------------------------------------------------------------------
// Equilibrium, evaluated for all Q populations from q-independent coefficients.
//   TruncationMa3 (Equilibrium.h:17-34): the reference's 9-term polynomial regrouped by powers of c.u
//     P = A0 + cu (A1 + cu (A2 + cu (A3 + cu A4))),  s = inv_cs2 (3 but for the multi-speed lattices)
//     A0 = 1 - s/2 u2 + s^2/8 u2^2, A1 = s - s^2/2 u2, A2 = s^2/2 - s^3/4 u2, A3 = s^3/6, A4 = s^4/24
//   Exact (Equilibrium.h:60-81, 106-126): product form, three factors per dimension precomputed.
// ------------------------------------------------------------------------------------------------
// CT: the arithmetic type -- double, or float for the BGK kernels on FP32 storage (the reference computes in its dataT)
template <class L, int EQ, typename CT = double> struct EquilibriumCoefficients;

// c_q . v for a compile-time celerity: additions and subtractions for the unit components, one multiply for the others
template <class L, int q, typename CT> __device__ __forceinline__ CT celerityDot(const CT* v) {
  CT r = (CT)0;
#pragma unroll
  for (int d = 0; d < L::D; ++d) {
    if (L::c(q, d) == 1) r += v[d];
    else if (L::c(q, d) == -1) r -= v[d];
    else if (L::c(q, d) != 0) r += (CT)L::c(q, d) * v[d];
  }
  return r;
}

template <class L, typename CT> struct EquilibriumCoefficients<L, kTruncationMa3, CT> {
  // s = L::inv_cs2: 3 for the single-speed lattices, where these constants are exactly 1.5, 1.125, 3, 4.5, 4.5, 6.75, 4.5, 3.375
  static constexpr double s = L::inv_cs2;
  static constexpr CT kA0u2 = (CT)(0.5 * s), kA0u4 = (CT)(0.125 * s * s), kA1 = (CT)s, kA1u2 = (CT)(0.5 * s * s), kA2 = (CT)(0.5 * s * s),
                      kA2u2 = (CT)(0.25 * s * s * s), kA3 = (CT)(s * s * s / 6.0), kA4 = (CT)(s * s * s * s / 24.0);
  CT a0, a1, a2;
  CT u[3];
  __device__ __forceinline__ void set(const CT* velocity, CT u2) {
    a0 = (CT)1 - kA0u2 * u2 + kA0u4 * u2 * u2;
    a1 = kA1 - kA1u2 * u2;
    a2 = kA2 - kA2u2 * u2;
#pragma unroll
    for (int d = 0; d < 3; ++d) u[d] = d < L::D ? velocity[d] : (CT)0;
  }
  // returns feq / (rho * w_q)
  template <int q> __device__ __forceinline__ CT shape() const {
    const CT cu = celerityDot<L, q, CT>(u);
    if (L::norm2(q) == 0) return a0;
    return a0 + cu * (a1 + cu * (a2 + cu * (kA3 + cu * kA4)));
  }
};

template <class L, typename CT> struct EquilibriumCoefficients<L, kExact, CT> {
  CT factor[3][3];  // [d][c+1]: (2 - sqrt(1+3u^2)) * ((2u + sqrt(1+3u^2)) / (1-u))^c
  __device__ __forceinline__ void set(const CT* velocity, CT) {
#pragma unroll
    for (int d = 0; d < L::D; ++d) {
      const CT ud = velocity[d];
      const CT root = sqrt((CT)1 + (CT)3 * ud * ud);
      const CT a = (CT)2 - root;
      const CT b = ((CT)2 * ud + root) / ((CT)1 - ud);
      factor[d][0] = a * ((CT)1 / b);
      factor[d][1] = a;
      factor[d][2] = a * b;
    }
  }
  template <int q> __device__ __forceinline__ CT shape() const {
    CT r = factor[0][L::c(q, 0) + 1];
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
// fastLogCore: natural logarithm of doubles in [0.25, 4) with an absolute error below 2.5e-16, 8 FP64 + 4 integer
// instructions instead of the ~30 FP64 + ~25 integer instructions of the CUDA math library's log().
// The entropic solve evaluates (1 + iterations) * Q logarithms per node, which makes it FP64-issue and shared-memory
// bound; this is what moves it back towards the HBM roofline.  The arguments are f_q / w_q and
// (f_q - alpha fNeq_q) / w_q, i.e. the local density times 1 + O(Mach) + O(non-equilibrium): [0.25, 4) covers every
// state a lattice-Boltzmann run can sensibly be in, and anything else takes the library path (entropicNewtonLibrary).
//   The top bits of the double (exponent and 8 mantissa bits, minus those of 0.25) index one of 1024 sub-intervals with
//   centre c; invc = 1/c rounded to a double with a ZERO LOW WORD, logc = double(-ln invc) (log_table.inc, generated with
//   100-digit arithmetic by gen_log_table.py); r = fma(v, invc, -1) is exact to rounding and |r| < 2^-9 + 2^-20, so
//   ln v = logc + (r - r^2/2 + ... + r^5/5) with a truncation error below 1e-17.
// Two table formats (LogTable<SPLIT>): {invc, logc} pairs, one 16-byte load per logarithm, or two arrays -- the high words of
// invc (4 bytes) and logc (8 bytes) -- 12 bytes and a third fewer shared-memory wavefronts, but one more load and two more
// integer instructions.  Measured (profiles/r02d_table_alphamax_variants.txt): the pairs win by 1-3 % on every lattice.
// N independent arguments advance in lock step: every Horner step is issued for all N before the next one, which
// gives the FP64 pipe N independent dependency chains per warp.
// `range` accumulates the maximum table index as an unsigned number: it stays below 1024 exactly when every argument
// was inside [0.25, 4) (smaller, negative, infinite and NaN arguments all map to indices >= 1024: the high word minus that
// of 0.25, in unsigned arithmetic, is either below 2^22 -- the table -- or at least 2^30).
// ------------------------------------------------------------------------------------------------
constexpr int kLogTableEntries = 1024;
// MLBM_LOG_TABLE_SPLIT = 1: two arrays, the high words of invc (4 bytes) and logc (8 bytes) -- 12 bytes and 3 shared-memory
// wavefronts per warp lookup, two loads; = 0: one array of {invc, logc} pairs -- 16 bytes and 4-5 wavefronts, one load
// and three integer instructions less.  Which one wins depends on whether a kernel is short of shared-memory bandwidth
// (D3Q27) or of issue slots (D2Q9): logTableSplit(Q).
#ifndef MLBM_LOG_TABLE_SPLIT_Q9
#define MLBM_LOG_TABLE_SPLIT_Q9 0
#endif
#ifndef MLBM_LOG_TABLE_SPLIT_Q27
#define MLBM_LOG_TABLE_SPLIT_Q27 0
#endif
constexpr bool logTableSplit(int Q) { return Q <= 13 ? MLBM_LOG_TABLE_SPLIT_Q9 != 0 : MLBM_LOG_TABLE_SPLIT_Q27 != 0; }
static __device__ const unsigned kLogInverseHigh[kLogTableEntries] = {
#define MLBM_LOG_ENTRY(inverseHigh, invc, logc) inverseHigh,
#include "log_table.inc"
#undef MLBM_LOG_ENTRY
};
static __device__ const double kLogCentre[kLogTableEntries] = {
#define MLBM_LOG_ENTRY(inverseHigh, invc, logc) logc,
#include "log_table.inc"
#undef MLBM_LOG_ENTRY
};
static __device__ const double2 kLogPairs[kLogTableEntries] = {
#define MLBM_LOG_ENTRY(inverseHigh, invc, logc) {invc, logc},
#include "log_table.inc"
#undef MLBM_LOG_ENTRY
};
template <bool SPLIT> struct LogTable;
template <> struct LogTable<true> {
  const unsigned* inverseHigh;  // high word of invc (its low word is zero)
  const double* logc;           // -ln(invc)
  static constexpr int kBytes = kLogTableEntries * 12;
  __device__ __forceinline__ void global() { inverseHigh = kLogInverseHigh; logc = kLogCentre; }
  // copies the table to `shared` (kBytes, 8-byte aligned) with all `threads` threads of the block; the caller synchronises
  __device__ __forceinline__ void stage(unsigned char* shared, int thread, int threads) {
    double* sharedLogc = reinterpret_cast<double*>(shared);
    unsigned* sharedHigh = reinterpret_cast<unsigned*>(sharedLogc + kLogTableEntries);
    for (int i = thread; i < kLogTableEntries; i += threads) { sharedLogc[i] = kLogCentre[i]; sharedHigh[i] = kLogInverseHigh[i]; }
    inverseHigh = sharedHigh;
    logc = sharedLogc;
  }
  __device__ __forceinline__ void fetch(unsigned index, double& invc, double& logCentre) const {
    invc = __hiloint2double((int)inverseHigh[index], 0);
    logCentre = logc[index];
  }
};
template <> struct LogTable<false> {
  const double2* pairs;  // {invc, -ln(invc)}
  static constexpr int kBytes = kLogTableEntries * 16;
  __device__ __forceinline__ void global() { pairs = kLogPairs; }
  __device__ __forceinline__ void stage(unsigned char* shared, int thread, int threads) {
    double2* sharedPairs = reinterpret_cast<double2*>(shared);
    for (int i = thread; i < kLogTableEntries; i += threads) sharedPairs[i] = kLogPairs[i];
    pairs = sharedPairs;
  }
  __device__ __forceinline__ void fetch(unsigned index, double& invc, double& logCentre) const {
    const double2 entry = pairs[index];
    invc = entry.x;
    logCentre = entry.y;
  }
};

template <int N, class Table>
__device__ __forceinline__ void fastLogCore(const double (&v)[N], double (&out)[N], const Table& table, unsigned& range) {
  double invc[N], logc[N], r[N], p[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const unsigned index = ((unsigned)__double2hiint(v[j]) - 0x3FD00000u) >> 12;  // unsigned: negative arguments wrap, they do not overflow
    range = max(range, index);
    table.fetch(index & (kLogTableEntries - 1), invc[j], logc[j]);
  }
#pragma unroll
  for (int j = 0; j < N; ++j) r[j] = fma(v[j], invc[j], -1.0);
#pragma unroll
  for (int j = 0; j < N; ++j) p[j] = fma(r[j], 0.2, -0.25);
#pragma unroll
  for (int j = 0; j < N; ++j) p[j] = fma(r[j], p[j], 1.0 / 3.0);
#pragma unroll
  for (int j = 0; j < N; ++j) p[j] = fma(r[j], p[j], -0.5);
#pragma unroll
  for (int j = 0; j < N; ++j) out[j] = logc[j] + fma(r[j] * r[j], p[j], r[j]);
}

// general entry point (self-test, tools): library logarithm outside [0.25, 4)
template <class Table>
__device__ __forceinline__ double fastLog(double v, const Table& table) {
  const double in[1] = {v};
  double out[1];
  unsigned range = 0;
  fastLogCore<1>(in, out, table, range);
  return range < kLogTableEntries ? out[0] : log(v);
}

// ------------------------------------------------------------------------------------------------
// Entropic alpha: Collision<ELBM>::calculateAlpha (Collision.h:351-375).
//
// Work layout.  The populations f_q and their non-equilibrium parts of the block's nodes are parked in SHARED memory
// ([row][node], conflict-free) while alpha is solved for.  That buys three things:
//   * the registers of the pull / moments / equilibrium phase are free during the solve;
//   * the solve is COMPACTED over the block: the nodes that left the small-deviation shortcut are listed in shared
//     memory and thread i solves the i-th listed node (any thread can read any node's column), so the FP64 cost follows
//     the number of such nodes instead of the number of warps that contain at least one of them;
//   * one block handles several x planes in a row, staging the logarithm table and the per-row constants once.
// Where the register file allows (columnRegisters(Q): both F2 and N2 for Q <= 13) the solving thread reads the column of its
// node ONCE into registers and unrolls the loops over q: the evaluations then cost only the table lookup in shared-memory
// traffic.  Larger lattices keep the rolled loops over the shared-memory column: with N2 (or both) in registers the D3Q27
// kernel needs 168 registers, spills, and loses 20 % (profiles/r02c_results.txt).  The kernels are bound by instruction
// issue, FP64 instructions counting double (profiles/r02_issue_model.md): every choice below is the one with fewer instructions.
//
// Arithmetic of one evaluation.  The populations of one speed class c (|c_q|^2 = 0, 1, 2, 3: same weight w_c) are
// stored next to each other and SCALED by 2^k_c, k_c the integer with w_c 2^k_c in [0.7, 1.4): the scaling is exact
// (it commutes with every rounding below), and with C_c = -ln(w_c 2^k_c)
//     ln(g_q / w_q) = ln(v_q) + C_c,    v_q = 2^k_c g_q = F2_q - a N2_q     (F2 = 2^k f, N2 = 2^k fNeq as stored),
// so the logarithm's argument comes out of ONE fused multiply-add, stays inside the table's [0.25, 4) for any
// sensible state, and the C_c terms collapse into per-node constants:
//   F(a)  = sum f ln(f/w) - g ln(g/w) = H - [sum_c 2^-k_c sum_{q in c} v_q L_q + A - a B]          (EntropicStep.h:31-45)
//   F'(a) = sum fNeq (1 + ln(g/w))    = sum fNeq + B + sum_c 2^-k_c sum_{q in c} N2_q L_q             (EntropicStep.h:47-62)
// with L_q = fastLog(v_q), A = sum_c 2^-k_c C_c sum_{q in c} F2_q, B likewise over N2, and H = F's first sum, hoisted
// out of the iteration: 11 FP64 instructions per population and evaluation (1 + 8 + 2), against ~3 x 60 as the
// reference writes it, and no per-population integer or constant traffic besides the table lookup.
// ------------------------------------------------------------------------------------------------
constexpr int logShift(double w) {
  int k = 0;
  while (w < 0.7) { w *= 2.0; ++k; }
  while (w >= 1.4) { w *= 0.5; --k; }
  return k;
}
constexpr double powerOfTwo(int k) {
  double value = 1.0;
  for (int i = 0; i < (k < 0 ? -k : k); ++i) value *= (k < 0 ? 0.5 : 2.0);
  return value;
}
// natural logarithm of y in [0.7, 1.4) at compile time: 2 atanh((y - 1) / (y + 1)), 40 terms (|s| < 0.18)
constexpr double constexprLog(double y) {
  const double s = (y - 1.0) / (y + 1.0), s2 = s * s;
  double term = s, sum = 0.0;
  for (int n = 1; n < 80; n += 2) { sum += term / n; term *= s2; }
  return 2.0 * sum;
}

// speed classes of a lattice: populations sorted by |c|^2 (stable), the order of the shared-memory rows
template <class L> struct SpeedClasses {
  static constexpr int count(int n2) {
    int n = 0;
    for (int q = 0; q < L::Q; ++q) n += L::norm2(q) == n2 ? 1 : 0;
    return n;
  }
  static constexpr int first(int n2) {
    int n = 0;
    for (int q = 0; q < L::Q; ++q) n += L::norm2(q) < n2 ? 1 : 0;
    return n;
  }
  static constexpr int row(int q) {
    int n = first(L::norm2(q));
    for (int other = 0; other < q; ++other) n += L::norm2(other) == L::norm2(q) ? 1 : 0;
    return n;
  }
  static constexpr double weight(int n2) { return detail::weightByNorm<L::id>(n2); }
  static constexpr int shift(int n2) { return logShift(weight(n2)); }
  static constexpr double scale(int n2) { return powerOfTwo(shift(n2)); }        // 2^k_c
  static constexpr double inverseScale(int n2) { return powerOfTwo(-shift(n2)); }  // 2^-k_c
  static constexpr double offset(int n2) { return -constexprLog(weight(n2) * scale(n2)); }  // C_c
  // logarithms advanced in lock step inside a class
  static constexpr int group(int n2) {
    const int n = count(n2);
#if defined(MLBM_LOG_GROUP_WIDE)
    return n % 6 == 0 ? 6 : (n % 8 == 0 ? 8 : (n % 4 == 0 ? 4 : (n % 3 == 0 ? 3 : (n < 3 ? (n > 0 ? n : 1) : 3))));
#else
    return n % 4 == 0 ? 4 : (n % 3 == 0 ? 3 : (n < 3 ? (n > 0 ? n : 1) : 3));
#endif
  }
};

template <int Q> struct EntropicShared {
  double* f;             // [Q][kStepBlock]  F2, rows sorted by speed class
  double* fNeq;          // [Q][kStepBlock]  N2
  double* alpha;         // [kStepBlock]     in: alphaMax of the nodes off the shortcut (when screened in registers), out: their alpha
  double* rowOffset;     // [kRowSlots]  C_c of the row's class       (library fallback only)
  double* rowInverse;    // [kRowSlots]  2^-k_c of the row's class    (library fallback only)
  int* warpCount;        // [kStepBlock / 32]
  unsigned char* list;   // [kStepBlock]     nodes (thread indices) that left the shortcut, ascending
  LogTable<logTableSplit(Q)> table;  // fastLog table (shared or global memory)
};

constexpr int logTableBytes(int Q) { return logTableSplit(Q) ? LogTable<true>::kBytes : LogTable<false>::kBytes; }
constexpr int rowSlots(int Q) { return Q <= 32 ? 32 : 40; }  // D3Q33 has 33 rows
constexpr int entropicSharedBytes(int Q, bool tableInShared) {
  return 2 * Q * kStepBlock * 8 + kStepBlock * 8 + 2 * rowSlots(Q) * 8 + 16 + kStepBlock + (tableInShared ? logTableBytes(Q) : 0);
}
// blocks per SM the entropic kernels are compiled for (registers) and sized for (shared memory)
// (measured, D3Q27 512^3: three blocks with the table in shared memory beat four blocks with the table in L1 by 10-14 %)
#ifndef MLBM_ENTROPIC_BLOCKS_Q9
#define MLBM_ENTROPIC_BLOCKS_Q9 5
#endif
constexpr int entropicBlocksPerSM(int Q) { return Q <= 9 ? MLBM_ENTROPIC_BLOCKS_Q9 : (Q >= 27 ? 3 : 4); }
// the fastLog table is staged in shared memory whenever those blocks still fit (228 KB, 1 KB reserved per block)
constexpr bool logTableInShared(int Q) { return entropicBlocksPerSM(Q) * (entropicSharedBytes(Q, true) + 1024) <= 233472; }
// which part of its node's column the solving thread keeps in registers: 2 = F2 and N2, 1 = N2, 0 = nothing (rolled loops)
#ifndef MLBM_COLUMN_REGISTERS_Q27
#define MLBM_COLUMN_REGISTERS_Q27 0
#endif
#ifndef MLBM_COLUMN_REGISTERS_Q19
#define MLBM_COLUMN_REGISTERS_Q19 0
#endif
// calculateAlphaMax (Collision.h:305-326) while fNeq is formed in registers (0: two multiplies and two compares per population
// on EVERY node) or by the thread that solves the node, from its column (1: only on the nodes that left the shortcut, but
// through shared memory when the column is not in registers)
#ifndef MLBM_ALPHAMAX_IN_SOLVER_Q9
#define MLBM_ALPHAMAX_IN_SOLVER_Q9 1
#endif
#ifndef MLBM_ALPHAMAX_IN_SOLVER_Q27
#define MLBM_ALPHAMAX_IN_SOLVER_Q27 0
#endif
constexpr bool alphaMaxInSolver(int Q) { return Q <= 13 ? MLBM_ALPHAMAX_IN_SOLVER_Q9 != 0 : MLBM_ALPHAMAX_IN_SOLVER_Q27 != 0; }
constexpr int columnRegisters(int Q) { return Q <= 13 ? 2 : (Q <= 21 ? MLBM_COLUMN_REGISTERS_Q19 : (Q <= 27 ? MLBM_COLUMN_REGISTERS_Q27 : 0)); }

// The same solve with the CUDA math library's logarithm, for the rare node whose arguments fall outside fastLogCore's
// table (and with it the reference's NaN behaviour for mirror states that leave the positive cone: the iteration produces
// NaNs, gives up after 50 iterations and alpha falls back to 2, EntropicStep.h:126-138, Collision.h:344-346).
// Kept out of line and rolled: it is cold code.
template <int Q>
__device__ __noinline__ double entropicNewtonLibrary(const double* fColumn, const double* nColumn, const double* rowOffset,
                                                     const double* rowInverse, double alphaGuess, double alphaMax) {
  double hoisted = 0.0;
#pragma unroll 1
  for (int row = 0; row < Q; ++row) {
    const double v = fColumn[row * kStepBlock];
    hoisted = fma(v * rowInverse[row], log(v) + rowOffset[row], hoisted);
  }
  double x = alphaGuess, step = 0.0;
  for (int iteration = 1; iteration <= 50; ++iteration) {
    x = x - step;
    double sum = 0.0, derivative = 0.0;
#pragma unroll 1
    for (int row = 0; row < Q; ++row) {
      const double n2 = nColumn[row * kStepBlock];
      const double v = fma(-x, n2, fColumn[row * kStepBlock]);
      const double lg = log(v) + rowOffset[row];
      sum = fma(v * rowInverse[row], lg, sum);
      derivative = fma(n2 * rowInverse[row], 1.0 + lg, derivative);
    }
    step = (hoisted - sum) / derivative;
    if (fabs(step) <= 1e-8) return (x > 1.0 && x < alphaMax) ? x : 2.0;
  }
  return 2.0;
}

// The column of one node as the solve sees it: MODE 2 keeps F2 and N2 in registers, MODE 1 N2 only (F2 is read from shared
// memory at compile-time offsets), MODE 0 nothing.  Rows are compile-time in MODE 1 / 2 (unrolled loops) and run-time in MODE 0.
template <int Q, int MODE> struct EntropicColumn {
  const double* fColumn;
  const double* nColumn;
  double fRegisters[MODE == 2 ? Q : 1];
  double nRegisters[MODE >= 1 ? Q : 1];
  __device__ __forceinline__ void load(const double* f, const double* n) {
    fColumn = f;
    nColumn = n;
    if constexpr (MODE == 2) {
#pragma unroll
      for (int row = 0; row < Q; ++row) fRegisters[row] = f[row * kStepBlock];
    }
    if constexpr (MODE >= 1) {
#pragma unroll
      for (int row = 0; row < Q; ++row) nRegisters[row] = n[row * kStepBlock];
    }
  }
  __device__ __forceinline__ double F(int row) const { if constexpr (MODE == 2) return fRegisters[row]; else return fColumn[row * kStepBlock]; }
  __device__ __forceinline__ double N(int row) const { if constexpr (MODE >= 1) return nRegisters[row]; else return nColumn[row * kStepBlock]; }
};

// one group of G rows: hC += v ln v, sF += F2, sN += N2
template <int G, int COUNT, class Column, class Table>
__device__ __forceinline__ void entropicHoistGroup(const Column& column, int firstRow, int inClass, const Table& table, unsigned& range,
                                                   double (&h)[G], double (&a)[G], double (&b)[G]) {
  double v[G], lg[G];
#pragma unroll
  for (int j = 0; j < G; ++j) {
    const bool live = COUNT % G == 0 || inClass + j < COUNT;
    v[j] = live ? column.F(firstRow + j) : 1.0;
    a[j] += live ? v[j] : 0.0;
    b[j] += live ? column.N(firstRow + j) : 0.0;
  }
  fastLogCore<G>(v, lg, table, range);
#pragma unroll
  for (int j = 0; j < G; ++j) h[j] = fma(v[j], lg[j], h[j]);
}

// one group of G rows of an evaluation: sum += v L, derivative += N2 L with v = F2 - x N2
template <int G, int COUNT, class Column, class Table>
__device__ __forceinline__ void entropicEvaluateGroup(const Column& column, int firstRow, int inClass, const Table& table, unsigned& range,
                                                      double x, double (&sum)[G], double (&derivative)[G]) {
  double n2[G], v[G], lg[G];
#pragma unroll
  for (int j = 0; j < G; ++j) {
    const bool live = COUNT % G == 0 || inClass + j < COUNT;
    n2[j] = live ? column.N(firstRow + j) : 0.0;
    v[j] = live ? fma(-x, n2[j], column.F(firstRow + j)) : 1.0;
  }
  fastLogCore<G>(v, lg, table, range);
#pragma unroll
  for (int j = 0; j < G; ++j) {
    sum[j] = fma(v[j], lg[j], sum[j]);
    derivative[j] = fma(n2[j], lg[j], derivative[j]);
  }
}

#ifndef MLBM_GROUP_UNROLL
#define MLBM_GROUP_UNROLL 1
#endif
constexpr int kGroupUnroll = MLBM_GROUP_UNROLL;  // groups of the rolled loops in flight per thread

// the groups of one speed class: unrolled when the column sits in registers (compile-time rows), rolled otherwise
template <int FIRST, int COUNT, int G, int MODE, class Body>
__device__ __forceinline__ void forEachGroup(Body&& body) {
  constexpr int groups = (COUNT + G - 1) / G;
  if constexpr (MODE >= 1) {
    staticFor<0, groups>([&](auto gc) { body(FIRST + decltype(gc)::value * G, decltype(gc)::value * G); });
  } else {
#pragma unroll kGroupUnroll
    for (int group = 0; group < groups; ++group) body(FIRST + group * G, group * G);
  }
}

// calculateAlpha for a node off the small-deviation shortcut (Collision.h:351-375; FORCED: :792-808), by the thread that
// solves it: calculateAlphaMax (Collision.h:305-326: min(2.5, min over fNeq_q > 0 of |f_q| / fNeq_q), tracked as a fraction;
// the 2^k scaling of a class cancels in the ratio), alphaMax < 2 -> 0.95 alphaMax, else solveAlpha (Collision.h:328-349) ->
// NewtonRaphsonSolver (EntropicStep.h:111-140), see the banner above.  `evaluations` returns the evaluations of (F, F').
template <class L>
__device__ __forceinline__ double entropicAlpha(const EntropicShared<L::Q>& s, int node, double alphaGuess, double alphaMaxScreened, int& evaluations) {
  using C = SpeedClasses<L>;
  constexpr int MODE = columnRegisters(L::Q);
  EntropicColumn<L::Q, MODE> column;
  column.load(s.f + node, s.fNeq + node);

  double alphaMax = alphaMaxScreened;
  if constexpr (alphaMaxInSolver(L::Q)) {
    double num = 2.5, den = 1.0;
    auto screen = [&](int row) {
      const double n2 = column.N(row);
      if (n2 > 0.0) {
        const double af = fabs(column.F(row));
        if (af * den < num * n2) { num = af; den = n2; }
      }
    };
    if constexpr (MODE >= 1) {
      staticFor<0, L::Q>([&](auto rc) { screen(decltype(rc)::value); });
    } else {
#pragma unroll 1
      for (int row = 0; row < L::Q; ++row) screen(row);
    }
    alphaMax = num / den;
  }
  evaluations = 0;
  if (alphaMax < 2.0) return 0.95 * alphaMax;

  unsigned range = 0;
  double hoisted = 0.0, offsetF = 0.0, offsetN = 0.0, sumN = 0.0;
  staticFor<0, L::maxNorm2() + 1>([&](auto nc) {
    constexpr int n2 = decltype(nc)::value;
    if constexpr (C::count(n2) > 0) {
      constexpr int G = C::group(n2);
      double h[G], a[G], b[G];
#pragma unroll
      for (int j = 0; j < G; ++j) h[j] = a[j] = b[j] = 0.0;
      forEachGroup<C::first(n2), C::count(n2), G, MODE>([&](int firstRow, int inClass) {
        entropicHoistGroup<G, C::count(n2)>(column, firstRow, inClass, s.table, range, h, a, b);
      });
      double hC = h[0], sF = a[0], sN = b[0];
#pragma unroll
      for (int j = 1; j < G; ++j) { hC += h[j]; sF += a[j]; sN += b[j]; }
      hoisted = fma(C::inverseScale(n2), hC, hoisted);
      offsetF = fma(C::inverseScale(n2) * C::offset(n2), sF, offsetF);
      offsetN = fma(C::inverseScale(n2) * C::offset(n2), sN, offsetN);
      sumN = fma(C::inverseScale(n2), sN, sumN);
    }
  });
  if (range >= kLogTableEntries) return entropicNewtonLibrary<L::Q>(column.fColumn, column.nColumn, s.rowOffset, s.rowInverse, alphaGuess, alphaMax);
  hoisted += offsetF;
  const double derivativeBase = sumN + offsetN;

  double x = alphaGuess, step = 0.0;
  bool converged = false;
  for (int iteration = 1; iteration <= 50; ++iteration) {
    x = x - step;
    evaluations = iteration;
    double total = 0.0, slope = 0.0;
    staticFor<0, L::maxNorm2() + 1>([&](auto nc) {
      constexpr int n2 = decltype(nc)::value;
      if constexpr (C::count(n2) > 0) {
        constexpr int G = C::group(n2);
        double sum[G], derivative[G];
#pragma unroll
        for (int j = 0; j < G; ++j) sum[j] = derivative[j] = 0.0;
        forEachGroup<C::first(n2), C::count(n2), G, MODE>([&](int firstRow, int inClass) {
          entropicEvaluateGroup<G, C::count(n2)>(column, firstRow, inClass, s.table, range, x, sum, derivative);
        });
        double sC = sum[0], dC = derivative[0];
#pragma unroll
        for (int j = 1; j < G; ++j) { sC += sum[j]; dC += derivative[j]; }
        total = fma(C::inverseScale(n2), sC, total);
        slope = fma(C::inverseScale(n2), dC, slope);
      }
    });
    if (range >= kLogTableEntries) return entropicNewtonLibrary<L::Q>(column.fColumn, column.nColumn, s.rowOffset, s.rowInverse, alphaGuess, alphaMax);
    step = (hoisted - (total + fma(-x, offsetN, offsetF))) / (derivativeBase + slope);
    if (fabs(step) <= 1e-8) { converged = (x > 1.0 && x < alphaMax); break; }
  }
  return converged ? x : 2.0;
}

// ------------------------------------------------------------------------------------------------
// Pieces shared by the two kernel bodies
// ------------------------------------------------------------------------------------------------
struct NodeIndex {
  int xh, m, r;  // the node itself: plane index counted from the first halo plane, row, column
  // element offsets INSIDE one population of the upstream plane / row / column for a celerity component +1, 0, -1 (pull from
  // x - c).  32-bit: a population (LX + 2 H planes) holds fewer than 2^32 elements (checked by mlbm_create), so that an
  // address costs one three-input add and one widening multiply-add instead of a chain of 64-bit operations.
  unsigned plane[3], row[3], column[3];
};

template <class L>
__device__ __forceinline__ NodeIndex nodeIndex(const StepParams& p, int x, int m, int r) {
  // upstream coordinates: pull from (x - cx, m - cm, r - cr) of the periodic image
  NodeIndex n;
  n.xh = x + L::H;  // L::H halo planes (one but for the multi-speed lattices) precede the interior
  n.m = m;
  n.r = r;
  int xPrev = n.xh - 1, xNext = n.xh + 1;
  if (p.wrapX) {
    if (xPrev == 0) xPrev = p.LX;
    if (xNext == p.LX + 1) xNext = 1;
  }
  const unsigned planeElements = (unsigned)p.plane, NR = (unsigned)p.NR;
  n.plane[0] = (unsigned)xPrev * planeElements;
  n.plane[1] = (unsigned)n.xh * planeElements;
  n.plane[2] = (unsigned)xNext * planeElements;
  n.row[0] = (unsigned)(m == 0 ? p.NM - 1 : m - 1) * NR;
  n.row[1] = (unsigned)m * NR;
  n.row[2] = (unsigned)(m == p.NM - 1 ? 0 : m + 1) * NR;
  n.column[0] = (unsigned)(r == 0 ? p.NR - 1 : r - 1);
  n.column[1] = (unsigned)r;
  n.column[2] = (unsigned)(r == p.NR - 1 ? 0 : r + 1);
  return n;
}

// periodic image of coordinate v - c in [0, n) for |c| up to 3 and any n >= 1
__device__ __forceinline__ int wrapCoordinate(int v, int n) {
  v %= n;
  return v < 0 ? v + n : v;
}

template <class L, typename StoreT, bool STREAMING = false, typename CT = double>
__device__ __forceinline__ void pullPopulations(const StepParams& p, const NodeIndex& n, CT (&f)[L::Q]) {
  const StoreT* __restrict__ prev = static_cast<const StoreT*>(p.prev);
  if constexpr (L::H > 1) {
    // multi-speed lattices (Lattice.h:213-458, 706-803): m and r wrap by index arithmetic whatever the length of the jump;
    // x wraps the same way on one rank and reaches into the L::H halo planes per side otherwise
    const int x = n.xh - L::H;
#pragma unroll
    for (int q = 0; q < L::Q; ++q) {
      const int xs = L::cx(q) == 0 ? n.xh : (p.wrapX ? wrapCoordinate(x - L::cx(q), p.LX) : x - L::cx(q)) + L::H;
      const int ms = L::cm(q) == 0 ? n.m : wrapCoordinate(n.m - L::cm(q), p.NM);
      const int rs = L::cr(q) == 0 ? n.r : wrapCoordinate(n.r - L::cr(q), p.NR);
      const StoreT* source = prev + q * p.stride + xs * p.plane + (long long)ms * p.NR + rs;
      f[q] = STREAMING ? loadPopulationStreaming<CT>(source) : loadPopulation<CT>(source);
    }
    return;
  }
#pragma unroll
  for (int q = 0; q < L::Q; ++q) {
    const unsigned offset = n.plane[1 - L::cx(q)] + n.row[1 - L::cm(q)] + n.column[1 - L::cr(q)];
    const StoreT* source = prev + q * p.stride + offset;
    f[q] = STREAMING ? loadPopulationStreaming<CT>(source) : loadPopulation<CT>(source);
  }
}

// element offset of the node itself inside one population of `next` (32-bit, see NodeIndex)
__device__ __forceinline__ unsigned ownOffset(const NodeIndex& n) { return n.plane[1] + n.row[1] + n.column[1]; }

// Moment::calculateDensity / calculateVelocity (Moment.h:14-47)
template <class L, typename CT>
__device__ __forceinline__ void moments(const CT (&f)[L::Q], CT& rho, CT& invRho, CT (&u)[3], CT& u2) {
  rho = f[0];
#pragma unroll
  for (int q = 1; q < L::Q; ++q) rho += f[q];
  u[0] = u[1] = u[2] = (CT)0;
#pragma unroll
  for (int q = 1; q < L::Q; ++q) {
#pragma unroll
    for (int d = 0; d < L::D; ++d) {
      if (L::c(q, d) == 1) u[d] += f[q];
      else if (L::c(q, d) == -1) u[d] -= f[q];
      else if (L::c(q, d) != 0) u[d] += (CT)L::c(q, d) * f[q];
    }
  }
  invRho = (CT)1 / rho;
  u2 = (CT)0;
#pragma unroll
  for (int d = 0; d < L::D; ++d) {
    u[d] *= invRho;
    u2 += u[d] * u[d];
  }
}

// Force::setForce at local interior coordinates (Collision.h:81-88): profiles precomputed on the host for the analytic
// forces, or Force<Generic>::setForce (Force.h:39-48) for the array-type forces: component iD of the force FIELD at the
// node's local index.  storeNodeFields writes the same values back on stored steps, like Algorithm::storeFields does.
template <class L, typename StoreT, typename CT>
__device__ __forceinline__ void bodyForce(const StepParams& p, int x, int m, int r, CT (&F)[3]) {
  F[0] = F[1] = F[2] = (CT)0;
  if (p.hasForce == 1) {
#pragma unroll
    for (int d = 0; d < L::D; ++d) {
      const int axis = p.forceAxis[d];
      if (axis >= 0) F[d] = (CT)__ldg(p.forceTable[d] + (axis == 0 ? x : (axis == 1 ? m : r)));
    }
  } else if (p.hasForce == 2) {
    const StoreT* field = static_cast<const StoreT*>(p.force) + ((long long)x * p.plane + (long long)m * p.NR + r);
#pragma unroll
    for (int d = 0; d < L::D; ++d) F[d] = (CT)field[d * p.fieldStride];
  }
}

// collision source term S_q (ForcingScheme.h:99-117 Guo, :184-197 ExactDifferenceMethod; None / ShanChen: 0)
template <class L, int EQ, int SCHEME, typename CT = double> struct SourceTerm {
  CT uF = (CT)0, guoFactor = (CT)0, rho = (CT)0;
  CT u[3], F[3];
  EquilibriumCoefficients<L, EQ, CT> shifted;  // EDM: feq at u + F / rho
  __device__ __forceinline__ void set(const StepParams& p, CT density, CT invRho, const CT (&velocity)[3], const CT (&force)[3]) {
    rho = density;
    guoFactor = (CT)p.guoFactor;
#pragma unroll
    for (int d = 0; d < 3; ++d) { u[d] = velocity[d]; F[d] = force[d]; }
    if (SCHEME == kSchemeGuo) {
#pragma unroll
      for (int d = 0; d < L::D; ++d) uF += u[d] * F[d];
    }
    if (SCHEME == kSchemeEDM) {
      CT v[3] = {(CT)0, (CT)0, (CT)0};
      CT v2 = (CT)0;
#pragma unroll
      for (int d = 0; d < L::D; ++d) {
        v[d] = u[d] + F[d] * invRho;
        v2 += v[d] * v[d];
      }
      shifted.set(v, v2);
    }
  }
  // feq is the equilibrium the scheme is handed: feq_q for BGK, f_q - fNeq_q for ELBM (Collision.h:252)
  template <int q> __device__ __forceinline__ CT value(CT feq) const {
    if (SCHEME == kSchemeGuo) {
      CT cF = (CT)0, cu = (CT)0;
#pragma unroll
      for (int d = 0; d < L::D; ++d) {
        if (L::c(q, d) == 1) { cF += F[d]; cu += u[d]; }
        else if (L::c(q, d) == -1) { cF -= F[d]; cu -= u[d]; }
        else if (L::c(q, d) != 0) { cF += (CT)L::c(q, d) * F[d]; cu += (CT)L::c(q, d) * u[d]; }
      }
      return guoFactor * (CT)L::w(q) * (cF - uF + (CT)L::inv_cs2 * cu * cF);
    }
    if (SCHEME == kSchemeEDM) return rho * (CT)L::w(q) * shifted.template shape<q>() - feq;
    return (CT)0;
  }
};

// Algorithm::storeFields (Algorithm.h:150-194) and the per-node terms of the scalar analyses (Analysis.h:53-61; summed in double)
template <class L, typename StoreT, typename CT>
__device__ __forceinline__ void storeNodeFields(const StepParams& p, long long node, CT rho, CT invRho, const CT (&u)[3],
                                                const CT (&F)[3], double& energy, double& speed2) {
  const bool fields = (p.isStored & 1) != 0;
  if (fields) static_cast<StoreT*>(p.density)[node] = (StoreT)rho;
  const CT half = p.hydroShift ? (CT)0.5 * invRho : (CT)0;
#pragma unroll
  for (int d = 0; d < L::D; ++d) {
    const CT v = u[d] + half * F[d];
    if (fields) {
      static_cast<StoreT*>(p.velocity)[d * p.fieldStride + node] = (StoreT)v;
      static_cast<StoreT*>(p.force)[d * p.fieldStride + node] = (StoreT)F[d];
    }
    energy += 0.5 * (double)rho * (double)v * (double)v;  // TotalEnergy (Analysis.h:53-61)
    speed2 += (double)v * (double)v;
  }
}

// block reduction of the observables: warp shuffles, then one value per warp through shared memory
__device__ __forceinline__ void reduceBlockObservables(const StepParams& p, int x, double energy, double mass, double speed2) {
#pragma unroll
  for (int offset = 16; offset > 0; offset >>= 1) {
    energy += __shfl_xor_sync(0xffffffffu, energy, offset);
    mass += __shfl_xor_sync(0xffffffffu, mass, offset);
    speed2 = fmax(speed2, __shfl_xor_sync(0xffffffffu, speed2, offset));
  }
  __shared__ double scratch[kObservableSlots][kStepBlock / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // the scratch array may still be read by the previous plane of this block
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
    const long long block = ((long long)x * p.NM + blockIdx.y) * gridDim.x + blockIdx.x;
    p.partials[block * kObservableSlots + 0] = e;
    p.partials[block * kObservableSlots + 1] = ms;
    p.partials[block * kObservableSlots + 2] = s2;
  }
}

// ------------------------------------------------------------------------------------------------
// Entropic body: Collision<ELBM> (Collision.h:182-376); see the banner above entropicAlpha.
// grid = (ceil(NR / kStepBlock), NM, ceil(planes / planesPerBlock)); the block walks planesPerBlock planes.
// ------------------------------------------------------------------------------------------------
//
// FORCED selects Collision<ForcedNR_ELBM_Forcing> (Collision.h:727-857): the populations handed to the solve are the
// forced ones, ff_q = f_q + S_q(feq_q) (calculateRelaxationTime :757-778), alphaMax is min |ff_q / fNeq_q| over
// fNeq_q > 0 (:810-832), there is no small-deviation shortcut (:792-808), the entropy condition is the mirror functor
// EntropicStepFunctor<T, true> (EntropicStep.h:65-108; the same F and F' with ff in place of f) and the collide is
// next = ff_q - alpha beta fNeq_q (:780-789).
template <class L, int EQ, int SCHEME, typename StoreT, bool FORCED>
__device__ __forceinline__ void entropicStepBody(const StepParams& p) {
  constexpr int Q = L::Q;
  using C = SpeedClasses<L>;
  extern __shared__ __align__(16) unsigned char dynamicShared[];
  EntropicShared<Q> s;
  s.f = reinterpret_cast<double*>(dynamicShared);
  s.fNeq = s.f + Q * kStepBlock;
  s.alpha = s.fNeq + Q * kStepBlock;
  s.rowOffset = s.alpha + kStepBlock;
  s.rowInverse = s.rowOffset + rowSlots(Q);
  s.warpCount = reinterpret_cast<int*>(s.rowInverse + rowSlots(Q));
  s.list = reinterpret_cast<unsigned char*>(s.warpCount + 4);
  s.table.global();
  staticFor<0, Q>([&](auto qc) {
    constexpr int q = decltype(qc)::value;
    if (threadIdx.x == q) {
      s.rowOffset[C::row(q)] = C::offset(L::norm2(q));
      s.rowInverse[C::row(q)] = C::inverseScale(L::norm2(q));
    }
  });
  if (logTableInShared(Q)) s.table.stage(dynamicShared + entropicSharedBytes(Q, false), threadIdx.x, kStepBlock);

  const int t = threadIdx.x;
  const int r = blockIdx.x * kStepBlock + t;
  const int m = blockIdx.y;
  const bool active = r < p.NR;
  double* const myF = s.f + t;
  double* const myN = s.fNeq + t;
  StoreT* __restrict__ next = static_cast<StoreT*>(p.next);
  StoreT* alphaField = static_cast<StoreT*>(p.alpha);

  for (int i = 0; i < p.planesPerBlock; ++i) {
    const int planeIndex = blockIdx.z * p.planesPerBlock + i;
    if (planeIndex >= p.planeCount) break;  // block-uniform
    const int x = p.x0 + planeIndex * p.planeStep;
    const long long rowNode = (long long)x * p.plane + (long long)m * p.NR;  // field / alpha index of r = 0
    __syncthreads();  // constants staged (first plane) / shared columns of the previous plane no longer read

    double rho = 0.0, invRho = 0.0, energy = 0.0, speed2 = 0.0, alpha = 2.0;
    double u[3] = {0.0, 0.0, 0.0}, F[3] = {0.0, 0.0, 0.0};
    bool offShortcut = false;
    if (active) {
      const NodeIndex n = nodeIndex<L>(p, x, m, r);
      double f[Q];
      pullPopulations<L, StoreT, !logTableInShared(Q)>(p, n, f);
      double u2;
      moments<L, double>(f, rho, invRho, u, u2);
      bodyForce<L, StoreT, double>(p, x, m, r, F);
      EquilibriumCoefficients<L, EQ> eq;
      eq.set(u, u2);
      // Collision<ELBM>::calculateRelaxationTime (Collision.h:227-241): fNeq, then alpha.  While fNeq is formed the cheap
      // screen of calculateAlpha runs on the register values: isDeviationSmall (Collision.h:284-303), no |fNeq_q| / f_q
      // above 1e-3.  Written as |N2_q| > 1e-3 F2_q, which is the reference's predicate for every F2_q >= 0 (the
      // division by zero included: inf > 1e-3 when fNeq_q != 0, NaN > 1e-3 false when it is 0); a node with a NEGATIVE
      // population -- sign bit of any F2_q, collected with integer ORs -- repeats the screen literally below.
      bool large = false;
      int signs = 0;
      double num = 2.5, den = 1.0;  // calculateAlphaMax as a fraction, when it is screened here (alphaMaxInSolver(Q) == false)
      auto trackAlphaMax = [&](double f2, double n2) {
        if constexpr (!alphaMaxInSolver(Q)) {
          if (n2 > 0.0) {
            const double af = fabs(f2);
            if (af * den < num * n2) { num = af; den = n2; }
          }
        }
      };
      if constexpr (FORCED) {
        SourceTerm<L, EQ, SCHEME> forcedSource;
        forcedSource.set(p, rho, invRho, u, F);
        staticFor<0, Q>([&](auto qc) {
          constexpr int q = decltype(qc)::value;
          constexpr double scale = C::scale(L::norm2(q));
          const double feq = rho * L::w(q) * eq.template shape<q>();
          // the population the entropy condition is written for: f_q + S_q
          const double f2 = (f[q] + forcedSource.template value<q>(feq)) * scale;  // exact
          const double n2 = (f[q] - feq) * scale;
          myF[C::row(q) * kStepBlock] = f2;
          myN[C::row(q) * kStepBlock] = n2;
          trackAlphaMax(f2, n2);
        });
        offShortcut = true;  // the forced variant has no isDeviationSmall shortcut
      } else {
        staticFor<0, Q>([&](auto qc) {
          constexpr int q = decltype(qc)::value;
          constexpr double scale = C::scale(L::norm2(q));
          // (f - feq) 2^k = f 2^k - feq 2^k and feq 2^k = (rho (w 2^k)) shape, all exact: one multiply less per population
          const double f2 = f[q] * scale;
          const double n2 = f2 - rho * (L::w(q) * scale) * eq.template shape<q>();
          myF[C::row(q) * kStepBlock] = f2;
          myN[C::row(q) * kStepBlock] = n2;
          large = large || fabs(n2) > 1.0e-3 * f2;
          signs |= __double2hiint(f2);
          trackAlphaMax(f2, n2);
        });
        if (signs < 0) {  // some population is negative (or -0): the reference's predicate, literally
          large = false;
#pragma unroll 1
          for (int row = 0; row < Q; ++row) {
            const double f2 = myF[row * kStepBlock], a = fabs(myN[row * kStepBlock]);
            large = large || (f2 > 0.0 ? (a > 1.0e-3 * f2) : (f2 == 0.0 ? a > 0.0 : false));
          }
        }
        offShortcut = large;
      }
      if constexpr (!alphaMaxInSolver(Q)) {
        if (offShortcut) s.alpha[t] = num / den;  // read by the thread that solves this node
      }
    }

    // compaction: the i-th node (in thread order) that left the shortcut is solved by thread i
    const unsigned ballot = __ballot_sync(0xffffffffu, offShortcut);
    const int warp = t >> 5, lane = t & 31;
    if (lane == 0) s.warpCount[warp] = __popc(ballot);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kStepBlock / 32; ++w) {
      const int count = s.warpCount[w];
      if (w < warp) before += count;
      total += count;
    }
    if (total > 0) {  // block-uniform
      if (offShortcut) s.list[before + __popc(ballot & ((1u << lane) - 1u))] = (unsigned char)t;
      __syncthreads();
      if (t < total) {
        const int node = s.list[t];
        const double guess = (double)alphaField[rowNode + blockIdx.x * kStepBlock + node];  // previous step's alpha (Algorithm.h:103-106)
        int evaluations = 0;
        s.alpha[node] = entropicAlpha<L>(s, node, guess, alphaMaxInSolver(Q) ? 0.0 : s.alpha[node], evaluations);
        // statistics of the solve (mlbm_newton_statistics: the FP64 side of the roofline); counted only while a caller asks
        if (p.newtonCounters && evaluations > 0) {
          atomicAdd(p.newtonCounters, 1ull);
          atomicAdd(p.newtonCounters + 1, (unsigned long long)evaluations);
        }
      }
      __syncthreads();
      if (offShortcut) alpha = s.alpha[t];
    }

    if (active) {
      const long long node = rowNode + r;
      const unsigned out = (unsigned)(x + L::H) * (unsigned)p.plane + (unsigned)m * (unsigned)p.NR + (unsigned)r;
      StoreT* const remoteHigh = (p.peerHigh && x == p.LX - 1) ? static_cast<StoreT*>(p.peerHigh) + ((long long)m * p.NR + r) : nullptr;
      StoreT* const remoteLow = (p.peerLow && x == 0) ? static_cast<StoreT*>(p.peerLow) + ((long long)(p.LX + 1) * p.plane + (long long)m * p.NR + r) : nullptr;
      alphaField[node] = (StoreT)alpha;
      const double omega = alpha * p.beta;  // 1 / tau_eff (Collision.h:240)
      SourceTerm<L, EQ, SCHEME> source;
      if (!FORCED) source.set(p, rho, invRho, u, F);
      // Collision<ELBM>::collideAndStream (Collision.h:243-258) / Collision<ForcedNR_ELBM_Forcing>::collideAndStream (:780-789,
      // the source is already inside the stored populations); the 2^k scaling of the stored values is undone exactly
      staticFor<0, Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        constexpr double inverse = C::inverseScale(L::norm2(q));
        const double f2 = myF[C::row(q) * kStepBlock], n2 = myN[C::row(q) * kStepBlock];
        double value = (f2 - omega * n2) * inverse;
        if (!FORCED) value += source.template value<q>((f2 - n2) * inverse);
        storePopulation(next + q * p.stride + out, value);
        if (L::cx(q) == 1 && remoteHigh) remoteHigh[q * p.stride] = (StoreT)value;
        if (L::cx(q) == -1 && remoteLow) remoteLow[q * p.stride] = (StoreT)value;
      });
      if (p.isStored) storeNodeFields<L, StoreT, double>(p, node, rho, invRho, u, F, energy, speed2);
    }
    if (p.isStored) reduceBlockObservables(p, x, energy, active ? rho : 0.0, speed2);
  }
}

// ------------------------------------------------------------------------------------------------
// The fused step.
// grid = (ceil(NR / kStepBlock), NM, number of x planes [/ planesPerBlock]), block = kStepBlock threads along r.
// ------------------------------------------------------------------------------------------------
template <class L, int COLLISION, int EQ, int SCHEME, typename StoreT>
#ifndef MLBM_BGK_BLOCKS
#define MLBM_BGK_BLOCKS 4
#endif
#ifndef MLBM_BGK_BLOCKS_F32
#define MLBM_BGK_BLOCKS_F32 4
#endif
__global__ void __launch_bounds__(kStepBlock, COLLISION != kBGK ? entropicBlocksPerSM(L::Q) : (sizeof(StoreT) == 4 ? MLBM_BGK_BLOCKS_F32 : MLBM_BGK_BLOCKS))
fusedStepKernel(const __grid_constant__ StepParams p) {
  if constexpr (COLLISION != kBGK) {
    entropicStepBody<L, EQ, SCHEME, StoreT, COLLISION == kELBMForcing>(p);
  } else {
    constexpr int Q = L::Q;
    // the reference computes in its dataT (Algorithm<T, ...>): FP32 storage means FP32 arithmetic here too, which is what
    // lets the FP32 kernels reach their (twice as high) HBM roofline instead of the FP64 pipe's; the observables' sums stay double
    using CT = StoreT;
    const int r = blockIdx.x * kStepBlock + threadIdx.x;
    const int m = blockIdx.y;
    const int x = p.x0 + (int)blockIdx.z * p.planeStep;
    const bool active = r < p.NR;
    double mass = 0.0, energy = 0.0, speed2 = 0.0;
    if (active) {
      const NodeIndex n = nodeIndex<L>(p, x, m, r);
      StoreT* __restrict__ next = static_cast<StoreT*>(p.next);
      CT f[Q];
      pullPopulations<L, StoreT, false, CT>(p, n, f);
      CT rho, invRho, u2, u[3], F[3];
      moments<L, CT>(f, rho, invRho, u, u2);
      bodyForce<L, StoreT, CT>(p, x, m, r, F);
      EquilibriumCoefficients<L, EQ, CT> eq;
      eq.set(u, u2);
      SourceTerm<L, EQ, SCHEME, CT> source;
      source.set(p, rho, invRho, u, F);

      const long long node = (long long)x * p.plane + (long long)m * p.NR + r;  // field index
      const unsigned out = ownOffset(n);
      // halo planes of the neighbours this node's outgoing populations belong to (block-uniform conditions)
      StoreT* const remoteHigh = (p.peerHigh && x == p.LX - 1) ? static_cast<StoreT*>(p.peerHigh) + ((long long)m * p.NR + r) : nullptr;
      StoreT* const remoteLow = (p.peerLow && x == 0) ? static_cast<StoreT*>(p.peerLow) + ((long long)(p.LX + 1) * p.plane + (long long)m * p.NR + r) : nullptr;
      // Collision<BGK>::collideAndStream (Collision.h:134-151)
      const CT keep = (CT)(1.0 - 2.0 * p.beta);
      const CT relax = (CT)(2.0 * p.beta);
      staticFor<0, Q>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        const CT feq = rho * (CT)L::w(q) * eq.template shape<q>();
        const CT value = keep * f[q] + relax * feq + source.template value<q>(feq);
        storePopulation(next + q * p.stride + out, value);
        if (L::cx(q) == 1 && remoteHigh) remoteHigh[q * p.stride] = (StoreT)value;
        if (L::cx(q) == -1 && remoteLow) remoteLow[q * p.stride] = (StoreT)value;
      });
      // BGK's alpha field is the constant 2 (Collision.h:121) and is not stored
      mass = (double)rho;
      if (p.isStored) storeNodeFields<L, StoreT, CT>(p, node, rho, invRho, u, F, energy, speed2);
    }
    if (p.isStored) reduceBlockObservables(p, x, energy, mass, speed2);
  }
}

using StepKernel = void (*)(const StepParams);

// Returns the specialised kernel for a configuration or nullptr when the reference has no such
// combination (e.g. the exact equilibrium exists only for D2Q9 and D3Q27, Equilibrium.h:36-126).
// Defined across the instantiate_*.cu translation units.
StepKernel lookupStepKernel(int lattice, int collision, int equilibrium, int scheme, int dtype);

}  // namespace mlbm
