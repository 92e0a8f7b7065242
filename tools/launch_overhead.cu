// Micro-probe: event-bracketed duration of back-to-back launches as a function of dynamic shared memory
// and of what runs between them (profiling aid for DESIGN.md's launch-overhead note).
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void spin_kernel(long long cycles) {
  extern __shared__ unsigned char smem[];
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
  if (cycles < 0) smem[threadIdx.x] = 0;
}
__global__ void small_kernel(int* p) { if (p && threadIdx.x == 1000) *p = 0; }
static float run(int smem, long long cycles, int mode, int iters) {
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaFuncSetAttribute(spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t* ev = new cudaEvent_t[2 * iters];
  for (int i = 0; i < 2 * iters; ++i) cudaEventCreate(&ev[i]);
  int* d; cudaMalloc(&d, 1 << 20);
  for (int w = 0; w < 2; ++w) {
    for (int i = 0; i < iters; ++i) {
      if (mode == 1) cudaMemsetAsync(d, 0, 1 << 18, st);
      if (mode == 2) small_kernel<<<64, 256, 0, st>>>(d);
      cudaEventRecord(ev[2 * i], st);
      spin_kernel<<<128, 320, smem, st>>>(cycles);
      cudaEventRecord(ev[2 * i + 1], st);
    }
    cudaStreamSynchronize(st);
  }
  float sum = 0;
  for (int i = 0; i < iters; ++i) { float ms; cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]); sum += ms; }
  return 1e3f * sum / iters;
}
static float run_batch(int smem, long long cycles, int iters, int threads, int blocks) {
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaFuncSetAttribute(spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int w = 0; w < 2; ++w) {
    cudaEventRecord(e0, st);
    for (int i = 0; i < iters; ++i) spin_kernel<<<blocks, threads, smem, st>>>(cycles);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  return 1e3f * ms / iters;
}
int main() {
  const int iters = 100;
  for (int smem : {0, 210 * 1024})
    for (int blocks : {128, 144, 148})
      printf("BATCH spin 26000 cycles, smem %3d KB, %d CTAs x 320 thr: %.2f us per launch (one event pair / 100 launches)\n",
             smem / 1024, blocks, run_batch(smem, 26000, iters, 320, blocks));
  const long long cyc[] = {0, 28000};
  for (long long c : cyc)
    for (int smem : {0, 64 * 1024, 210 * 1024})
      for (int mode : {0, 1, 2})
        printf("spin %6lld cycles, smem %3d KB, between=%s : %.2f us per launch (event-bracketed)\n", c, smem / 1024,
               mode == 0 ? "nothing" : (mode == 1 ? "memset " : "kernel "), run(smem, c, mode, iters));
  return 0;
}
