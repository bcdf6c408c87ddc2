// tests/emu/include/cuda_host_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// The slice of the CUDA runtime API that csrc/context.cu, shell_force.cu and spectral.cu call, implemented on the host so
// that the library's HOST logic (allocation, pitched copies, the per-step orchestration, the observables pipeline) can be
// executed by the CPU test-suite around the emulated kernels (tests/emu/cuda_emu.h): device memory is host memory, streams
// execute immediately, events are time stamps.  Device allocations live in POSIX shared memory so that another rank (process)
// can map them through the cudaIpc* calls, which is how the direct peer halos work on the box.  See tests/emu/build_context.py.
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorNotSupported = 801 };
typedef struct CUevent_emu { double stamp; }* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocMapped = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaIpcMemHandle_t { char reserved[64]; };

namespace cuda_emu {
inline cudaError_t& lastError() { static cudaError_t e = cudaSuccess; return e; }
inline double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace cuda_emu

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorMemoryAllocation ? "out of memory" : "emulated CUDA error"); }
inline cudaError_t cudaGetLastError() { const cudaError_t e = cuda_emu::lastError(); cuda_emu::lastError() = cudaSuccess; return e; }
inline cudaError_t cudaGetDeviceCount(int* count) { *count = 8; return cudaSuccess; }   // one emulated device per rank (process)
inline cudaError_t cudaSetDevice(int device) { return device >= 0 && device < 8 ? cudaSuccess : cudaErrorInvalidValue; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* least, int* greatest) { *least = 0; *greatest = -5; return cudaSuccess; }

namespace cuda_emu {
struct Allocation { std::string name; size_t bytes; };
inline std::map<void*, Allocation>& allocations() { static auto* m = new std::map<void*, Allocation>(); return *m; }
inline void* mapSegment(const char* name, size_t bytes, bool create) {
  const int fd = shm_open(name, create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
  if (fd < 0) return nullptr;
  if (create && ftruncate(fd, (off_t)bytes) != 0) { close(fd); shm_unlink(name); return nullptr; }
  void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  return p == MAP_FAILED ? nullptr : p;
}
}  // namespace cuda_emu

namespace cuda_emu {
inline void* allocate(size_t bytes) {
#ifdef MLBM_EMU_ASAN
  // AddressSanitizer build (MLBM_EMULATED_FLAGS="-fsanitize=address -DMLBM_EMU_ASAN"): plain heap blocks, so that an index
  // error of a kernel or of the host code lands in a red zone (single-rank runs only: nothing can be mapped by a peer)
  void* heap = std::malloc(bytes ? bytes : 1);
  if (heap) { std::memset(heap, 0xFF, bytes); allocations()[heap] = Allocation{"", bytes ? bytes : 1}; }
  return heap;
#endif
  static int counter = 0;
  char name[64];
  std::snprintf(name, sizeof(name), "/mlbm_emu_mem_%d_%d", (int)getpid(), counter++);
  const size_t size = bytes ? bytes : 1;
  void* p = mapSegment(name, size, true);
  if (!p) return nullptr;
  // poisoned like fresh device memory is not: reads of bytes nobody wrote show up as NaNs / huge integers
  std::memset(p, 0xFF, size);
  allocations()[p] = Allocation{name, size};
  return p;
}
}  // namespace cuda_emu

template <class T> inline cudaError_t cudaMalloc(T** pointer, size_t bytes) {
  void* p = cuda_emu::allocate(bytes);
  if (!p) return cudaErrorMemoryAllocation;
  *pointer = static_cast<T*>(p);
  return cudaSuccess;
}
inline cudaError_t cudaFree(void* pointer) {
  if (!pointer) return cudaSuccess;
  auto found = cuda_emu::allocations().find(pointer);
  if (found == cuda_emu::allocations().end()) return cudaErrorInvalidValue;
#ifdef MLBM_EMU_ASAN
  std::free(pointer);
#else
  munmap(pointer, found->second.bytes);
  shm_unlink(found->second.name.c_str());
#endif
  cuda_emu::allocations().erase(found);
  return cudaSuccess;
}
template <class T> inline cudaError_t cudaMallocHost(T** pointer, size_t bytes) { *pointer = static_cast<T*>(std::malloc(bytes ? bytes : 1)); return *pointer ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> inline cudaError_t cudaHostAlloc(T** pointer, size_t bytes, unsigned) { return cudaMallocHost(pointer, bytes); }
inline cudaError_t cudaFreeHost(void* pointer) { std::free(pointer); return cudaSuccess; }
inline cudaError_t cudaMemset(void* p, int value, size_t bytes) { std::memset(p, value, bytes); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int value, size_t bytes, cudaStream_t = nullptr) { std::memset(p, value, bytes); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) { std::memmove(dst, src, bytes); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t = nullptr) { std::memmove(dst, src, bytes); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind,
                                     cudaStream_t = nullptr) {
  if (width > dpitch || width > spitch) return cudaErrorInvalidValue;   // what the runtime rejects as "invalid pitch"
  for (size_t row = 0; row < height; ++row) std::memcpy(static_cast<char*>(dst) + row * dpitch, static_cast<const char*>(src) + row * spitch, width);
  return cudaSuccess;
}

inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* stream, unsigned, int) { *stream = std::malloc(1); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t stream) { std::free(stream); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* event) { *event = new CUevent_emu{0.0}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* event, unsigned) { return cudaEventCreate(event); }
inline cudaError_t cudaEventDestroy(cudaEvent_t event) { delete event; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t event, cudaStream_t = nullptr) { event->stamp = cuda_emu::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t start, cudaEvent_t stop) { *ms = (float)(stop->stamp - start->stamp); return cudaSuccess; }
inline cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int bytes) { return bytes <= 227 * 1024 ? cudaSuccess : cudaErrorInvalidValue; }
// CUDA IPC: the handle names the shared-memory segment behind a cudaMalloc of another process
struct cudaIpcHandleEmu { char name[48]; unsigned long long bytes; };
static_assert(sizeof(cudaIpcHandleEmu) <= sizeof(cudaIpcMemHandle_t), "handle size");
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* handle, void* pointer) {
  auto found = cuda_emu::allocations().find(pointer);
  if (found == cuda_emu::allocations().end()) return cudaErrorInvalidValue;
  cudaIpcHandleEmu h = {};
  std::snprintf(h.name, sizeof(h.name), "%s", found->second.name.c_str());
  h.bytes = found->second.bytes;
  std::memset(handle, 0, sizeof(*handle));
  std::memcpy(handle, &h, sizeof(h));
  return cudaSuccess;
}
namespace cuda_emu { inline std::map<void*, size_t>& mappings() { static auto* m = new std::map<void*, size_t>(); return *m; } }
inline cudaError_t cudaIpcOpenMemHandle(void** pointer, cudaIpcMemHandle_t handle, unsigned) {
  cudaIpcHandleEmu h;
  std::memcpy(&h, &handle, sizeof(h));
  void* p = cuda_emu::mapSegment(h.name, (size_t)h.bytes, false);
  if (!p) return cudaErrorInvalidValue;
  cuda_emu::mappings()[p] = (size_t)h.bytes;
  *pointer = p;
  return cudaSuccess;
}
inline cudaError_t cudaIpcCloseMemHandle(void* pointer) {
  auto found = cuda_emu::mappings().find(pointer);
  if (found == cuda_emu::mappings().end()) return cudaErrorInvalidValue;
  munmap(pointer, found->second);
  cuda_emu::mappings().erase(found);
  return cudaSuccess;
}
