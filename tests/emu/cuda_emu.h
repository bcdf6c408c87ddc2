// tests/emu/cuda_emu.h -- TEST INFRASTRUCTURE ONLY (never compiled into, loaded by or shipped with metalbm_b200/).
//
// A minimal host-side stand-in for the CUDA execution model, just large enough to run the SOURCE of the fused step
// kernels (metalbm_b200/csrc/step_kernel.cuh) on the CPU so that the CPU test-suite can check their LOGIC -- index
// arithmetic, the plane loop, the shared-memory phases and barriers of the entropic kernel, the block-level compaction,
// the peer halo stores -- against the oracle without a GPU.  It says nothing about performance, memory coalescing or
// hardware rounding (host libm / host FMA), and the product never falls back to it.
//
// Model: one block at a time; every CUDA thread of the block is a ucontext fiber; __syncthreads() and the warp
// collectives (__ballot_sync, __shfl_xor_sync) yield to a round-robin scheduler until all participants have arrived.
// Shared memory is poisoned with NaN bit patterns before every block so that reads of unwritten shared memory show.
#pragma once

#include <sched.h>
#include <ucontext.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__
#define __constant__
#define __shared__ static
#define __align__(n) alignas(n)

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 { double x, y; };
typedef void* cudaStream_t;

namespace cuda_emu {

constexpr int kMaxThreads = 1024;
constexpr size_t kStackBytes = 256 * 1024;
constexpr size_t kSharedBytes = 256 * 1024;

struct Fiber {
  ucontext_t context;
  std::vector<unsigned char> stack;
  bool done = false;
  uint3 index{0, 0, 0};
};

struct State {
  ucontext_t scheduler;
  std::vector<Fiber> fibers;
  int current = -1;
  int blockThreads = 0;
  // block barrier
  int barrierArrived = 0;
  unsigned long long barrierGeneration = 0;
  // per-warp collective exchange
  struct Warp {
    double values[32];
    unsigned long long bits[32];
    int arrived = 0, departed = 0;
    unsigned long long generation = 0;
  } warps[kMaxThreads / 32];
  unsigned long long progress = 0;
  alignas(16) unsigned char shared[kSharedBytes];
};

inline State& state() { static State* s = new State(); return *s; }
inline unsigned char* dynamicSharedBase() { return state().shared; }

}  // namespace cuda_emu

// the built-in variables: plain globals, set by the scheduler before a fiber resumes
inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

namespace cuda_emu {

inline void yield() {
  State& s = state();
  const int me = s.current;
  swapcontext(&s.fibers[me].context, &s.scheduler);
}

// all participants of a warp-wide exchange deposit, wait for the others, read, and wait again before the slots are reused
template <class Read>
inline auto warpCollective(double value, unsigned long long bits, Read&& read) {
  State& s = state();
  const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
  State::Warp& w = s.warps[warp];
  const int lanes = std::min(32, s.blockThreads - warp * 32);
  while (w.departed != 0) yield();  // previous collective of this warp still being read
  w.values[lane] = value;
  w.bits[lane] = bits;
  ++w.arrived;
  ++s.progress;
  while (w.arrived < lanes) yield();
  auto result = read(w, lane, lanes);
  if (++w.departed == lanes) { w.arrived = 0; w.departed = 0; }
  ++s.progress;
  return result;
}

inline void blockBarrier() {
  State& s = state();
  const unsigned long long generation = s.barrierGeneration;
  ++s.progress;
  if (++s.barrierArrived == s.blockThreads) {
    s.barrierArrived = 0;
    ++s.barrierGeneration;
    return;
  }
  while (s.barrierGeneration == generation) yield();
}

// what every fiber of the running launch executes (set by the launch functions below)
inline std::function<void()>& thunk() { static std::function<void()> f; return f; }

inline void fiberEntry() {
  thunk()();
  State& s = state();
  s.fibers[s.current].done = true;
  ++s.progress;
  swapcontext(&s.fibers[s.current].context, &s.scheduler);
}

inline unsigned scheduleOrder(unsigned slot, unsigned block) {
  static const int mode = [] {
    const char* e = std::getenv("MLBM_EMU_SCHEDULE");
    return !e ? 0 : (std::strcmp(e, "reverse") == 0 ? 1 : (std::strcmp(e, "random") == 0 ? 2 : 0));
  }();
  if (mode == 1) return block - 1 - slot;
  if (mode == 2) {  // a fresh pseudo-random rotation + stride per round (coprime stride: a permutation)
    static unsigned long long state = 0x9E3779B97F4A7C15ull;
    static unsigned offset = 0, stride = 1, lastBlock = 0;
    if (slot == 0 || lastBlock != block) {
      state = state * 6364136223846793005ull + 1442695040888963407ull;
      offset = (unsigned)(state >> 33) % block;
      do { state = state * 6364136223846793005ull + 1442695040888963407ull; stride = (unsigned)(state >> 33) % block; } while (stride == 0 || std::__gcd(stride, block) != 1);
      lastBlock = block;
    }
    return (offset + slot * stride) % block;
  }
  return slot;
}

// one block after the other; every thread of a block is a fiber
inline void runGrid(dim3 grid, unsigned block, size_t sharedBytes) {
  State& s = state();
  if (block > (unsigned)kMaxThreads || sharedBytes > kSharedBytes) { std::fprintf(stderr, "cuda_emu: launch too large\n"); std::abort(); }
  gridDim = grid;
  blockDim = dim3(block, 1, 1);
  s.blockThreads = (int)block;
  if (s.fibers.size() < block) s.fibers.resize(block);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::memset(s.shared, 0xFF, sizeof(s.shared));  // NaN poison
        s.barrierArrived = 0;
        for (auto& w : s.warps) { w.arrived = 0; w.departed = 0; }
        for (unsigned t = 0; t < block; ++t) {
          Fiber& f = s.fibers[t];
          if (f.stack.empty()) f.stack.resize(kStackBytes);
          f.done = false;
          f.index = uint3{t, 0, 0};
          getcontext(&f.context);
          f.context.uc_stack.ss_sp = f.stack.data();
          f.context.uc_stack.ss_size = f.stack.size();
          f.context.uc_link = &s.scheduler;
          makecontext(&f.context, (void (*)())fiberEntry, 0);
        }
        unsigned remaining = block;
        while (remaining > 0) {
          const unsigned long long before = s.progress;
          remaining = 0;
          for (unsigned slot = 0; slot < block; ++slot) {
            // MLBM_EMU_SCHEDULE=reverse|random: other legal interleavings of the block's threads -- a kernel that misses a
            // barrier computes something else under them (a poor man's race detector)
            const unsigned t = scheduleOrder(slot, block);
            Fiber& f = s.fibers[t];
            if (f.done) continue;
            s.current = (int)t;
            threadIdx = f.index;
            blockIdx = uint3{bx, by, bz};
            swapcontext(&s.scheduler, &f.context);
            if (!f.done) ++remaining;
          }
          if (remaining > 0 && s.progress == before) {
            std::fprintf(stderr, "cuda_emu: deadlock in block (%u, %u, %u): %u threads wait for a barrier the others never reach\n", bx, by, bz, remaining);
            std::abort();
          }
        }
      }
}

// kernel<<<grid, block, sharedBytes>>>(params) for the one-struct step kernels
template <class Params>
inline void launch(void (*kernel)(const Params), dim3 grid, unsigned block, size_t sharedBytes, const Params& params) {
  thunk() = [kernel, &params]() { kernel(params); };
  runGrid(grid, block, sharedBytes);
}

// kernel<<<grid, block, sharedBytes, stream>>>(args...) for any kernel (tests/emu/build_context.py rewrites the launches of
// csrc/context.cu, shell_force.cu and spectral.cu into calls of this); streams are synchronous here
template <class... Params, class... Args>
inline void launchKernel(void (*kernel)(Params...), dim3 grid, dim3 block, size_t sharedBytes, Args&&... args) {
  std::tuple<std::decay_t<Params>...> stored(std::forward<Args>(args)...);
  thunk() = [kernel, &stored]() { std::apply(kernel, stored); };
  runGrid(grid, block.x * block.y * block.z, sharedBytes);
}

}  // namespace cuda_emu

// ---- intrinsics used by the kernels --------------------------------------------------------------------------
inline void __syncthreads() { cuda_emu::blockBarrier(); }

inline unsigned __ballot_sync(unsigned, bool predicate) {
  return cuda_emu::warpCollective(0.0, predicate ? 1ull : 0ull, [](cuda_emu::State::Warp& w, int, int lanes) {
    unsigned mask = 0;
    for (int i = 0; i < lanes; ++i) mask |= (unsigned)(w.bits[i] & 1ull) << i;
    return mask;
  });
}

inline double __shfl_xor_sync(unsigned, double value, int laneMask) {
  return cuda_emu::warpCollective(value, 0ull, [laneMask](cuda_emu::State::Warp& w, int lane, int lanes) {
    const int source = lane ^ laneMask;
    return source < lanes ? w.values[source] : w.values[lane];
  });
}

inline int __popc(unsigned value) { return __builtin_popcount(value); }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline void __stcs(T* p, T value) { *p = value; }
inline int __double2hiint(double value) { uint64_t bits; std::memcpy(&bits, &value, 8); return (int)(bits >> 32); }
inline int __double2loint(double value) { uint64_t bits; std::memcpy(&bits, &value, 8); return (int)(bits & 0xffffffffu); }
inline double __hiloint2double(int hi, int lo) {
  const uint64_t bits = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double value;
  std::memcpy(&value, &bits, 8);
  return value;
}
using std::fabs;
using std::fma;
using std::fmax;
using std::log;
using std::max;
using std::min;
using std::sqrt;

// ---- intrinsics used by the kernels of csrc/context.cu, shell_force.cu and spectral.cu -----------------------------
inline unsigned atomicAdd(unsigned* address, unsigned value) { const unsigned old = *address; *address = old + value; return old; }
inline unsigned long long atomicAdd(unsigned long long* address, unsigned long long value) { const unsigned long long old = *address; *address = old + value; return old; }
inline double atomicAdd(double* address, double value) { const double old = *address; *address = old + value; return old; }
inline void __threadfence() {}
inline void __threadfence_system() {}
inline long long clock64() {   // nanoseconds ~ cycles at 1 GHz: the peer-flag wait gives up after tens of seconds, as on the box
  return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline void __nanosleep(unsigned) { ++cuda_emu::state().progress; sched_yield(); cuda_emu::yield(); }   // waits on another rank (process): not a deadlock
inline void sincospi(double x, double* s, double* c) { *s = std::sin(M_PI * x); *c = std::cos(M_PI * x); }
inline double sinpi(double x) { return std::sin(M_PI * x); }
