// Micro-probe: latency and throughput of FP64 arithmetic / conversions on this part (design input for the
// association kernel's exact path and the Kalman kernels; results quoted in profiles/README.md).
#include <cuda_runtime.h>
#include <stdio.h>
template <int MODE>
__global__ void probe(double* out, double seed, int iters, long long* cycles) {
  double a = seed + threadIdx.x, b = 1.0000001, c = 0.5;
  double a2 = a + 1, a3 = a + 2, a4 = a + 3;
  float f = (float)seed + threadIdx.x;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) { a = fma(a, b, c); }                                     // dependent DFMA chain
    if (MODE == 1) { a = fma(a, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c); }  // 4 chains
    if (MODE == 2) { a = floor(a * b) + c; }                                 // floor
    if (MODE == 3) { a = a / b + c; }                                        // division
    if (MODE == 4) { f = fmaf(f, 1.0000001f, 0.5f); }                        // FFMA chain (reference)
    if (MODE == 5) { a = (double)(int)(a) + c; }                             // F2I / I2F fp64
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + a2 + a3 + a4 + f;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMallocManaged(&cyc, 8);
  const char* names[] = {"DFMA dependent chain", "DFMA 4 independent chains", "floor(fp64)+mul+add", "fp64 divide+add", "FFMA dependent chain", "F2I+I2F fp64 + add"};
  const int iters = 2000;
  for (int warps : {1, 8, 32}) {
    printf("--- %d warp(s) per SM, 148 CTAs ---\n", warps);
#define RUN(M) { probe<M><<<148, 32 * warps>>>(out, 1.5, iters, cyc); cudaDeviceSynchronize(); \
      printf("%-28s %8.1f cycles/iteration (per warp)\n", names[M], (double)*cyc / iters); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
  }
  return 0;
}
