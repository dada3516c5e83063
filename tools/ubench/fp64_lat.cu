// Single-warp FP64 latency / issue micro-benchmarks (clock64 around unrolled chains).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_lat fp64_lat.cu && ./fp64_lat
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
template <int ILP>
__global__ void dfma_chain(double* out, long long* cyc, double a, double b) {
  double v[ILP];
  for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x + i;
  long long t0 = clock64();
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += v[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP>
__global__ void op_chain(double* out, long long* cyc, double a) {
  double v = 1.5 + threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll
  for (int k = 0; k < 64; ++k) {
    if (OP == 0) v = sqrt(v) + a;
    if (OP == 1) v = a / v + 1.0;
    if (OP == 2) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v)); v = y + a; }
    if (OP == 3) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v)); v = y + a; }
    if (OP == 4) v = v + a;
    if (OP == 5) v = v * a;
    if (OP == 6) v = (v > a) ? v * a : v + a;
    if (OP == 7) v = __shfl_xor_sync(0xffffffffu, v, 1) + a;
    if (OP == 8) v = pow(v, -0.2) + a;
  }
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// 4 independent sqrt chains interleaved by hand vs the compiler's IEEE sqrt
__global__ void sqrt4(double* out, long long* cyc, double a) {
  double v[4];
  for (int i = 0; i < 4; ++i) v[i] = 1.5 + threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
#pragma unroll
  for (int k = 0; k < 32; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = sqrt(v[i]) + a;
  long long t1 = clock64();
  out[threadIdx.x] = v[0] + v[1] + v[2] + v[3];
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__device__ __forceinline__ double sqrt_sl(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  return fma(fma(-g, g, x), h, g);
}
template <int ILP>
__global__ void sqrt_sl_k(double* out, long long* cyc, double a) {
  double v[ILP];
  for (int i = 0; i < ILP; ++i) v[i] = 1.5 + threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
#pragma unroll
  for (int k = 0; k < 32; ++k)
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = sqrt_sl(v[i]) + a;
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += v[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
  auto rep = [&](const char* n, double per) { cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-44s %8lld cycles  %7.1f per op\n", n, h, h / per); };
  for (int threads : {32, 2}) {
    printf("-- %d active lanes, one warp\n", threads);
    for (int w = 0; w < 2; ++w) dfma_chain<1><<<1, threads>>>(out, cyc, 1.0000001, 1e-9); rep("DFMA dependent chain", N);
    for (int w = 0; w < 2; ++w) dfma_chain<2><<<1, threads>>>(out, cyc, 1.0000001, 1e-9); rep("DFMA 2 chains (per DFMA)", 2 * N);
    for (int w = 0; w < 2; ++w) dfma_chain<4><<<1, threads>>>(out, cyc, 1.0000001, 1e-9); rep("DFMA 4 chains (per DFMA)", 4 * N);
    for (int w = 0; w < 2; ++w) dfma_chain<8><<<1, threads>>>(out, cyc, 1.0000001, 1e-9); rep("DFMA 8 chains (per DFMA)", 8 * N);
    for (int w = 0; w < 2; ++w) dfma_chain<16><<<1, threads>>>(out, cyc, 1.0000001, 1e-9); rep("DFMA 16 chains (per DFMA)", 16 * N);
    for (int w = 0; w < 2; ++w) op_chain<4><<<1, threads>>>(out, cyc, 1e-9); rep("DADD dependent", 64);
    for (int w = 0; w < 2; ++w) op_chain<5><<<1, threads>>>(out, cyc, 1.0000001); rep("DMUL dependent", 64);
    for (int w = 0; w < 2; ++w) op_chain<6><<<1, threads>>>(out, cyc, 1.0000001); rep("DSETP+select+op dependent", 64);
    for (int w = 0; w < 2; ++w) op_chain<0><<<1, threads>>>(out, cyc, 1.0); rep("sqrt() + DADD dependent", 64);
    for (int w = 0; w < 2; ++w) op_chain<1><<<1, threads>>>(out, cyc, 1.3); rep("div + DADD dependent", 64);
    for (int w = 0; w < 2; ++w) op_chain<2><<<1, threads>>>(out, cyc, 1.0); rep("rsqrt.approx.f64 + DADD dependent", 64);
    for (int w = 0; w < 2; ++w) op_chain<3><<<1, threads>>>(out, cyc, 1.0); rep("rcp.approx.f64 + DADD dependent", 64);
    for (int w = 0; w < 2; ++w) op_chain<7><<<1, threads>>>(out, cyc, 1.0); rep("shfl(double) + DADD dependent", 64);
    for (int w = 0; w < 2; ++w) op_chain<8><<<1, threads>>>(out, cyc, 1.0); rep("pow + DADD dependent", 64);
    for (int w = 0; w < 2; ++w) sqrt4<<<1, threads>>>(out, cyc, 1.0); rep("sqrt() x4 independent (per sqrt)", 128);
    for (int w = 0; w < 2; ++w) sqrt_sl_k<1><<<1, threads>>>(out, cyc, 1.0); rep("straight-line sqrt dependent", 32);
    for (int w = 0; w < 2; ++w) sqrt_sl_k<4><<<1, threads>>>(out, cyc, 1.0); rep("straight-line sqrt x4 independent (per sqrt)", 128);
  }
  return 0;
}
