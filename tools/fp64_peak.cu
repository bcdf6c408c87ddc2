// fp64_peak.cu -- measures the FP64 (DFMA) issue peak of the GPU, the SECOND roofline of the entropic kernels
// ("ELBM additionally has an FP64-pipe bound -- report both").  MEASURED_PEAKS.json (driver-written) has no
// FP64 figure, so this tool writes its own file next to it: FP64_PEAK.json.
//
//   tools/_build/fp64_peak [device] [out.json]
//
// Every thread advances 8 independent fused multiply-add chains (enough to cover the DFMA latency at any occupancy);
// the grid is 8 blocks of 256 threads per SM; time from CUDA events, best of 10.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

constexpr int kChains = 8;
constexpr int kInner = 4096;

__global__ void __launch_bounds__(256) dfmaKernel(double* out, double a, double b) {
  double x[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) x[i] = 1.0 + 1e-9 * (threadIdx.x + i);
#pragma unroll 1
  for (int n = 0; n < kInner; ++n) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) x[i] = fma(x[i], a, b);
  }
  double sum = 0.0;
#pragma unroll
  for (int i = 0; i < kChains; ++i) sum += x[i];
  if (sum == 12345.678) out[0] = sum;  // never true: keeps the chains alive
}

int main(int argc, char** argv) {
  const int device = argc > 1 ? atoi(argv[1]) : 0;
  const char* path = argc > 2 ? argv[2] : nullptr;
  if (cudaSetDevice(device) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  double* out;
  cudaMalloc(&out, 8);
  const int blocks = prop.multiProcessorCount * 8;
  cudaEvent_t start, stop;
  cudaEventCreate(&start);
  cudaEventCreate(&stop);
  for (int i = 0; i < 3; ++i) dfmaKernel<<<blocks, 256>>>(out, 0.999999, 1e-7);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int repeat = 0; repeat < 10; ++repeat) {
    cudaEventRecord(start);
    dfmaKernel<<<blocks, 256>>>(out, 0.999999, 1e-7);
    cudaEventRecord(stop);
    cudaEventSynchronize(stop);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, start, stop);
    if (ms < best) best = ms;
  }
  if (cudaGetLastError() != cudaSuccess) { fprintf(stderr, "launch failed\n"); return 1; }
  const double laneOps = (double)blocks * 256 * kChains * kInner;  // DFMA lane-operations
  const double perSecond = laneOps / (best * 1e-3);
  int clockKHz = 0;
  cudaDeviceGetAttribute(&clockKHz, cudaDevAttrClockRate, device);
  char line[1024];
  snprintf(line, sizeof(line),
           "{\"gpu_name\": \"%s\", \"sms\": %d, \"dfma_lane_ops_per_s\": %.6e, \"fp64_tflops\": %.3f, "
           "\"dfma_per_clock_per_sm_at_max_clock\": %.2f, \"sm_max_mhz\": %.1f, \"kernel_ms\": %.4f, "
           "\"how\": \"tools/fp64_peak.cu: 8 independent DFMA chains per thread, 8 x 256 threads per SM, CUDA events, best of 10\"}\n",
           prop.name, prop.multiProcessorCount, perSecond, 2.0 * perSecond / 1e12,
           perSecond / prop.multiProcessorCount / (clockKHz * 1e3), clockKHz / 1e3, best);
  fputs(line, stdout);
  if (path) {
    FILE* file = fopen(path, "w");
    if (file) { fputs(line, file); fclose(file); }
  }
  cudaFree(out);
  return 0;
}
