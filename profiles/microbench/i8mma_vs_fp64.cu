// i8mma_vs_fp64.cu — do tcgen05.mma kind::i8 (UTCIMMA) and the FP64 pipe run concurrently on B200?
// 8 warps per SM run independent DFMA chains; optionally one more thread issues back-to-back int8 MMAs
// (M = 128, N = 32 / 64 / 128 / 256, K = 32; A from shared memory or tensor memory).  Reported: the DFMA rate
// with and without the tensor work, and the time per MMA with and without the DFMA work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o i8mma_vs_fp64 i8mma_vs_fp64.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);
}

// fp_iters: DFMA loop length per FP64 warp (0 = no FP64 work); mma_count: MMAs issued by the tensor thread (0 = none)
__global__ void __launch_bounds__(320, 1) k(double* out, int fp_iters, int mma_count, int N, int a_tmem, long long* cyc, int R) {
  extern __shared__ __align__(128) unsigned char smem[];  // A: 4 KB, B: 8 KB (zeros are fine)
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 12288 / 4; e += 320) reinterpret_cast<uint32_t*>(smem)[e] = 0x01010101u * (e & 3);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp < 8) {
    double a0 = tid * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < fp_iters; i++) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * 256 + tid] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  } else if ((tid == 256 || (tid == 288 && R < 0)) && mma_count > 0) {
    const int second = tid == 288;  // R < 0: two issuing threads (warps 8 and 9), |R| accumulators each, own barrier
    const int RR = R < 0 ? -R : R;
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da = make_desc(smem_u32(smem), 128, 256), db = make_desc(smem_u32(smem + 4096), 128, 256);
    const long long t0 = clock64();
    for (int i = 0; i < mma_count; i++) {
      const uint32_t d = tmem + (uint32_t)(((i & (RR - 1)) + second * RR) * N);  // RR (power of two) independent accumulators per issuing thread
      if (a_tmem)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
                     "r"(tmem + 504), "l"(db), "r"(idesc), "r"(1u)
                     : "memory");
      else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                     "l"(da), "l"(db), "r"(idesc), "r"(1u)
                     : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[second])) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok)
                   : "r"(smem_u32(&bars[second])), "r"(0u)
                   : "memory");
    if (blockIdx.x == 0 && !second) cyc[0] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 256 * 8);
  cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto run = [&](int fp_iters, int mma_count, int N, int a_tmem, const char* name, int R = 2) {
    float best = 1e30f;
    long long hc = 0;
    for (int rep = 0; rep < 3; rep++) {
      cudaMemset(cyc, 0, 8);
      cudaEventRecord(e0);
      k<<<148, 320, 16384>>>(out, fp_iters, mma_count, N, a_tmem, cyc, R);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
      cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
    }
    const double tf = 148.0 * 256 * 8.0 * fp_iters * 2 / (best * 1e-3) / 1e12;
    printf("%-46s %8.3f ms  DFMA %6.2f TF/s   %8.1f clk/MMA (%lld clk for %d MMAs, %.0f int8 MAC/clk/SM)\n", name, best, tf,
           mma_count ? (double)hc / mma_count / (R < 0 ? 2 : 1) : 0.0, hc, mma_count * (R < 0 ? 2 : 1), mma_count ? 128.0 * N * 32 * mma_count * (R < 0 ? 2 : 1) / (double)hc : 0.0);
  };
  const int F = 200000;
  run(F, 0, 32, 0, "DFMA only");
  for (int a_tmem = 0; a_tmem < 2; a_tmem++)
    for (int N = 32; N <= 256; N *= 2) {
      char nm[96];
      snprintf(nm, sizeof nm, "MMA only, N=%d, A from %s", N, a_tmem ? "TMEM" : "SMEM");
      run(0, 40000, N, a_tmem, nm);
    }
  for (int N : {32, 64, 96, 128, 192})
    for (int R : {1, 2, 4, 8, -1, -2}) {
      const int RR = R < 0 ? -2 * R : R;
      if (RR * N > 496) continue;
      char nm[96];
      snprintf(nm, sizeof nm, "MMA only, N=%d, A TMEM, %d accumulators x %d issuing thread(s)", N, R < 0 ? -R : R, R < 0 ? 2 : 1);
      run(0, 40000, N, 1, nm, R);
    }
  for (int a_tmem = 0; a_tmem < 2; a_tmem++)
    for (int N = 32; N <= 256; N *= 4) {
      char nm[96];
      snprintf(nm, sizeof nm, "DFMA + MMA N=%d, A from %s", N, a_tmem ? "TMEM" : "SMEM");
      const int cnt = N == 32 ? 60000 : 20000;
      run(F, cnt, N, a_tmem, nm);
    }
  return 0;
}
