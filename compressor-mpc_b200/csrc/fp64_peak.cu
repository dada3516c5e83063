// FP64 peak micro-benchmarks for the roofline denominator (the
// control step is bound by the FP64 pipe, and MEASURED_PEAKS.json holds no FP64
// figure, so it is measured live on the box that runs the bench).
//   - DFMA: 8 independent fused multiply-add chains per thread
//   - DMMA: mma.sync m8n8k4 / m16n8k8 f64, 4 independent accumulator sets per warp
// Exposed through the C ABI as cmpc_measure_fp64_peak (include/cmpc.h).
#include <cuda_runtime.h>

#include <cstdio>

#include "cmpc.h"

namespace {

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  double x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) dmma884_kernel(double* out, int iters, double a, double b) {
  double c[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j] = threadIdx.x + j;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dmma884(c[0], c[1], a, b);
      dmma884(c[2], c[3], a, b);
      dmma884(c[4], c[5], a, b);
      dmma884(c[6], c[7], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

__global__ void __launch_bounds__(256) dmma1688_kernel(double* out, int iters, double av, double bv) {
  double c[4][4];
  double a[4] = {av, av + 1, av + 2, av + 3}, b[2] = {bv, bv + 1};
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < 4; ++k) c[j][k] = threadIdx.x + j + k;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dmma1688(c[0], a, b);
      dmma1688(c[1], a, b);
      dmma1688(c[2], a, b);
      dmma1688(c[3], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < 4; ++k) s += c[j][k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double time_best_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;
}

}  // namespace

extern "C" int cmpc_measure_fp64_peak(int device, cmpc_fp64_peak* out) {
  if (!out) return CMPC_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return CMPC_ERR_CUDA;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  const int sms = prop.multiProcessorCount;
  const int blocks = sms * 8, threads = 256, iters = 2000;
  double* buf = nullptr;
  if (cudaMalloc(&buf, sizeof(double) * blocks * threads) != cudaSuccess) return CMPC_ERR_CUDA;
  const double total_threads = double(blocks) * threads;
  // DFMA: iters * 8 * 8 FMAs per thread
  double ms = time_best_ms([&] { dfma_kernel<<<blocks, threads>>>(buf, iters, 0.999999, 1e-9); }, 5);
  out->dfma_tflops = total_threads * iters * 64.0 * 2.0 / (ms * 1e-3) / 1e12;
  // DMMA m8n8k4: 16 mma per iter per warp, 2*8*8*4 flops each
  ms = time_best_ms([&] { dmma884_kernel<<<blocks, threads>>>(buf, iters, 0.999999, 1e-9); }, 5);
  out->dmma_m8n8k4_tflops = (total_threads / 32) * iters * 16.0 * 512.0 / (ms * 1e-3) / 1e12;
  // DMMA m16n8k8: 16 mma per iter per warp, 2*16*8*8 flops each
  ms = time_best_ms([&] { dmma1688_kernel<<<blocks, threads>>>(buf, iters, 0.999999, 1e-9); }, 5);
  out->dmma_m16n8k8_tflops = (total_threads / 32) * iters * 16.0 * 2048.0 / (ms * 1e-3) / 1e12;
  out->sm_count = sms;
  cudaFree(buf);
  return cudaGetLastError() == cudaSuccess ? CMPC_OK : CMPC_ERR_CUDA;
}
