// fp64_pipes.cu — microbenchmark: FP64 DFMA vs DMMA (mma.sync.m8n8k4.f64) throughput on B200,
// alone and mixed, to decide whether the X.Theta contraction of tiled_loglik should move to DMMA
// while the transcendental epilogue stays on the DFMA pipe.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>  // 0 dfma, 1 dmma, 2 mixed in the same warp (1 dmma : 8 dfma), 3 mixed by warp parity
__global__ void __launch_bounds__(256) k(double* out, double a, double b, int iters) {
  double x[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { x[i] = threadIdx.x * 1e-9 + i; c[i] = i * 1e-3; }
  const bool dm = (MODE == 1) || (MODE == 3 && ((threadIdx.x >> 5) & 1));
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (MODE == 0 || (MODE == 3 && !dm)) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
      } else if (MODE == 1 || (MODE == 3 && dm)) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) dmma(c[i], c[i + 1], a, b);
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
        dmma(c[0], c[1], a, b);
        dmma(c[2], c[3], a, b);
        dmma(c[4], c[5], a, b);
        dmma(c[6], c[7], a, b);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i] + c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double fma_per_thread_iter, double mma_per_warp_iter) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int blocks = sms * 8, threads = 256, iters = 4096;
  double* d; cudaMalloc(&d, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, 0.999999, 1e-7, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r) best = ms < best ? ms : best;
  }
  double threads_total = (double)blocks * threads;
  double dfma_flops = 2.0 * fma_per_thread_iter * iters * threads_total;
  double dmma_flops = 2.0 * 256.0 * mma_per_warp_iter * iters * (threads_total / 32);
  printf("%-28s %8.3f ms  DFMA %7.2f TF/s  DMMA %7.2f TF/s  total %7.2f TF/s\n", name, best,
         dfma_flops / best / 1e9, dmma_flops / best / 1e9, (dfma_flops + dmma_flops) / best / 1e9);
  cudaFree(d);
}

int main() {
  run<0>("dfma only", 64, 0);
  run<1>("dmma only", 0, 32);
  run<2>("mixed same warp 64f:32m", 64, 32);
  run<3>("mixed by warp (half/half)", 32, 16);
  return 0;
}
