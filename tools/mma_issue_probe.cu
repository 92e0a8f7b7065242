// tools/mma_issue_probe.cu -- how many SM cycles one tcgen05.mma (kind::f16, fp16 x fp16 -> fp32 in TMEM,
// M = 128 per CTA, K = 16) costs on a B200 as a function of N, for cta_group::1 and cta_group::2, when a
// single thread issues them back to back from shared-memory operands (128-byte swizzled K-major tiles, the
// layout TMA writes) -- the issue pattern of the association kernel's main loop
// (bot-sort-onnx-tensorrt_b200/csrc/reid_gemm.cu).  Operand contents are irrelevant (zeros).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_issue_probe tools/mma_issue_probe.cu
//   tools/bin/mma_issue_probe > profiles/r02_mma_issue_probe.txt
//
// Output: one line per (cta_group, N): cycles per MMA, the dense rate that corresponds to (flop / cycle / SM),
// and the fraction of the 8192 flop/cycle/SM datapath (128 x 256-wide: M128 N256 K16 in 128 cycles).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ uint64_t kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr int kStages = 4;          // operand slots cycled through, like the kernel's TMA ring
constexpr int kABytes = 128 * 64 * 2;
constexpr int kBBytes = 256 * 64 * 2;

// kGroup = 1: one CTA, M = 128.  kGroup = 2: a CTA pair (cluster of 2), M = 256 across the pair, the leader issues;
// each CTA holds its own 128 rows of A and HALF of the N columns of B (N / 2 rows of the B tile).
template <int kGroup>
__global__ void __launch_bounds__(128, 1) probe_kernel(int N, int n_mma, long long* out_all) {
  long long* out_cycles = out_all + 4 * (blockIdx.x / kGroup);
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kStages * (kABytes + kBBytes) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    if (kGroup == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (kGroup == 2) cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_ptr;
  const bool leader = kGroup == 1 || cluster_rank() == 0;
  if (leader && warp == 0 && lane == 0) {
    // instruction descriptor: D fp32 (1<<4), A/B fp16 K-major, N>>3 at [17,23), M>>4 at [24,29)
    const int M = kGroup == 1 ? 128 : 256;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const int stage = (i >> 2) % kStages, kk = i & 3;
      const uint32_t sa = smem_u32(smem + stage * (kABytes + kBBytes));
      const uint64_t adesc = kmajor_sw128_desc(sa) + (uint64_t)(kk * 2);
      const uint64_t bdesc = kmajor_sw128_desc(sa + kABytes) + (uint64_t)(kk * 2);
      const uint32_t accumulate = i != 0;
      if (kGroup == 1) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_base), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
      } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_base), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
      }
    }
    const long long t_issue = clock64();
    if (kGroup == 1)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    out_cycles[0] = t1 - t0;
    out_cycles[1] = t_issue - t0;
    out_cycles[2] = (long long)(g1 - g0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (kGroup == 2) cluster_sync();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (kGroup == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}


// The association kernel's ring protocol without any operand traffic: a producer thread waits for a free
// stage and marks it full at once; the issuer waits for the full stage, issues `per_stage` MMAs and commits
// the stage's "empty" barrier.  What does the barrier round trip cost the tensor pipe per stage?
//   wait_ahead = 0: wait(full[s]) right before the stage's MMAs (the obvious loop)
//   wait_ahead = 1: the NEXT stage's full barrier is waited for before this stage's commit is issued
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__global__ void __launch_bounds__(128, 1) ring_kernel(int N, int n_stage_rounds, int per_stage, int wait_ahead, int fence,
                                                      long long* out_all) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], done_bar;
  __shared__ uint32_t tmem_ptr;
  long long* out = out_all + 4 * blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kStages * (kABytes + kBBytes) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_ptr;
  if (warp == 1 && lane == 0) {           // producer
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < n_stage_rounds; ++i) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      mbar_arrive(&full_bar[stage]);
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  }
  if (warp == 0 && lane == 0) {           // issuer
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int stage = 0; uint32_t phase = 0;
    const long long t0 = clock64();
    if (wait_ahead) mbar_wait(&full_bar[0], 0);
    for (int i = 0; i < n_stage_rounds; ++i) {
      if (!wait_ahead) mbar_wait(&full_bar[stage], phase);
      if (fence) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = smem_u32(smem + stage * (kABytes + kBBytes));
      for (int kk = 0; kk < per_stage; ++kk) {
        const uint64_t adesc = kmajor_sw128_desc(sa) + (uint64_t)((kk & 3) * 2);
        const uint64_t bdesc = kmajor_sw128_desc(sa + kABytes) + (uint64_t)((kk & 3) * 2);
        const uint32_t accumulate = (i | kk) != 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_base), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
      }
      const int cur = stage;
      if (++stage == kStages) { stage = 0; phase ^= 1; }
      if (wait_ahead && i + 1 < n_stage_rounds) mbar_wait(&full_bar[stage], phase);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty_bar[cur])) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
    mbar_wait(&done_bar, 0);
    out[0] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}
static void run_ring(int N, int per_stage, int wait_ahead, int fence, int ctas, long long* d_out) {
  const size_t smem = (size_t)kStages * (kABytes + kBBytes) + 1024;
  CK(cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int rounds = 4096 / per_stage;
  long long best = -1;
  for (int rep = 0; rep < 5; ++rep) {
    ring_kernel<<<ctas, 128, smem>>>(N, rounds, per_stage, wait_ahead, fence, d_out);
    CK(cudaDeviceSynchronize());
    long long h[4 * 256];
    CK(cudaMemcpy(h, d_out, 32 * ctas, cudaMemcpyDeviceToHost));
    long long worst = 0;
    for (int g = 0; g < ctas; ++g) if (h[4 * g] > worst) worst = h[4 * g];
    if (best < 0 || worst < best) best = worst;
  }
  printf("ring  N=%3d  %2d MMAs per stage  wait_ahead=%d fence=%d  %3d SMs  %8.2f cycles/MMA  (%7.1f cycles per stage round, %6.1f over the issue bound)\n",
         N, per_stage, wait_ahead, fence, ctas, (double)best / 4096, (double)best / rounds, (double)best / rounds - per_stage * N / 2.0);
}

template <int kGroup>
static void run(int N, int n_mma, long long* d_out, int groups = 1) {
  const size_t smem = (size_t)kStages * (kABytes + kBBytes) + 1024;
  CK(cudaFuncSetAttribute(probe_kernel<kGroup>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(kGroup * groups); lc.blockDim = dim3(128); lc.dynamicSmemBytes = smem; lc.stream = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kGroup; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  long long best = -1, best_issue = 0, best_ns = 0;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaMemset(d_out, 0, 32 * 256));
    CK(cudaLaunchKernelEx(&lc, probe_kernel<kGroup>, N, n_mma, d_out));
    CK(cudaDeviceSynchronize());
    long long h[4 * 256];
    CK(cudaMemcpy(h, d_out, 32 * groups, cudaMemcpyDeviceToHost));
    // with several CTAs: the SLOWEST CTA of the launch (all SMs issue at the same time), best of 5 launches
    long long worst = 0, worst_issue = 0, worst_ns = 0;
    for (int g = 0; g < groups; ++g) if (h[4 * g] > worst) { worst = h[4 * g]; worst_issue = h[4 * g + 1]; worst_ns = h[4 * g + 2]; }
    if (best < 0 || worst < best) { best = worst; best_issue = worst_issue; best_ns = worst_ns; }
  }
  const double cyc = (double)best / n_mma;
  const double M = kGroup == 1 ? 128.0 : 256.0;
  const double flop_per_cyc_sm = 2.0 * M * N * 16.0 / cyc / kGroup;
  if (groups == 1)
    printf("cta_group::%d  M=%3d N=%3d K=16  %8.2f cycles/MMA (issue loop alone %7.2f)  %8.1f flop/cycle/SM  %5.1f %% of 8192\n", kGroup,
           (int)M, N, cyc, (double)best_issue / n_mma, flop_per_cyc_sm, 100.0 * flop_per_cyc_sm / 8192.0);
  else
    printf("cta_group::%d  M=%3d N=%3d K=16  %3d SMs busy  %8.2f cycles/MMA  %7.2f ns/MMA (slowest CTA)  %8.1f flop/cycle/SM  %5.1f %% of 8192  -> %6.0f TFLOP/s over the busy SMs\n",
           kGroup, (int)M, N, groups * kGroup, cyc, (double)best_ns / n_mma, flop_per_cyc_sm, 100.0 * flop_per_cyc_sm / 8192.0,
           2.0 * M * N * 16.0 * groups / ((double)best_ns / n_mma) / 1e3);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("# %s, %d SMs, SM clock (max) %d MHz; %d back-to-back tcgen05.mma kind::f16 per measurement, best of 5\n", prop.name,
         prop.multiProcessorCount, clk / 1000, 2048);
  printf("# peak at 8192 flop/cycle/SM x %d SMs x %.3f GHz = %.0f TFLOP/s dense fp16\n", prop.multiProcessorCount, clk / 1e6,
         8192.0 * prop.multiProcessorCount * clk / 1e9);
  long long* d_out;
  CK(cudaMalloc(&d_out, 32 * 256));
  const int Ns[] = {64, 96, 112, 128, 160, 192, 208, 224, 240, 256};
  for (int N : Ns) run<1>(N, 2048, d_out);
  for (int N : Ns) if (N % 32 == 0 || N % 16 == 0) run<2>(N, 2048, d_out);   // cta_group::2: N in steps of 16 (32 for M=256 kind::f16 needs N % 16 == 0)
  // the same issue loop on many SMs at once: is the per-SM rate a chip-wide quantity?
  printf("# all CTAs of a launch issue at the same time; slowest CTA, best of 5 launches, 8192 MMAs each\n");
  for (int N : {128, 224, 256})
    for (int g : {1, 18, 37, 74, 111, 148}) run<1>(N, 8192, d_out, g);
  for (int N : {224, 256})
    for (int g : {1, 37, 74}) run<2>(N, 8192, d_out, g);
  printf("# ring protocol of the association kernel (4 stages, producer thread marks a freed stage full at once), no operand traffic\n");
  for (int N : {224, 256})
    for (int per : {4, 8, 16})
      for (int wa : {0, 1})
        for (int fe : {1, 0}) run_ring(N, per, wa, fe, 1, d_out);
  run_ring(224, 4, 0, 1, 148, d_out);
  run_ring(256, 4, 0, 1, 148, d_out);
  CK(cudaFree(d_out));
  return 0;
}
