// fp64_issue.cu — microbenchmark: does the B200 SM keep its FP64 pipe (64 DFMA/clk/SM) full while
// the same warps also issue integer / FP32 / shared-memory instructions?  The tiled_loglik main loop
// issues ~0.56 non-FP64 instructions per FP64 instruction at 8 warps/SM; this measures how much DFMA
// throughput such a mix can retain, at the kernel's occupancy (8 warps/SM) and at full occupancy.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_issue fp64_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

// NI integer ops (IMAD / LOP3, independent chains) per 8 DFMAs; KIND 0 = IMAD, 1 = FFMA, 2 = LDS.64 broadcast
template <int NI, int KIND, int ILP>
__global__ void __launch_bounds__(256) k(double* out, double a, double b, int iters, int q) {
  __shared__ double sh[256];
  sh[threadIdx.x] = threadIdx.x;
  __syncthreads();
  double x[ILP];
  unsigned y[8];
  float f[8];
  double acc = 0.0;
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
  for (int i = 0; i < 8; i++) { y[i] = threadIdx.x + i; f[i] = threadIdx.x * 1e-3f + i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
      // NI extra instructions per ILP DFMAs
#pragma unroll
      for (int i = 0; i < NI; i++) {
        if (KIND == 0) y[i & 7] = y[i & 7] * 3u + (unsigned)q;
        else if (KIND == 1) f[i & 7] = fmaf(f[i & 7], 0.999f, 1e-3f);
        else acc += sh[(it + u + i * 8) & 255];  // uniform address -> broadcast LDS (+1 DADD, counted)
      }
    }
  }
  double s = acc;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
#pragma unroll
  for (int i = 0; i < 8; i++) s += (double)y[i] + (double)f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NI, int KIND, int ILP>
void run(const char* name, int blocks_per_sm) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int blocks = sms * blocks_per_sm, threads = 256, iters = 8192 / blocks_per_sm * 2;
  double* d; cudaMalloc(&d, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0);
    k<NI, KIND, ILP><<<blocks, threads>>>(d, 0.999999, 1e-7, iters, 7);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r) best = ms < best ? ms : best;
  }
  double nd = (double)ILP + (KIND == 2 ? NI : 0);
  double flops = 2.0 * nd * 8 * iters * (double)blocks * threads;
  printf("%-44s warps/SM %2d  %8.3f ms  FP64 %6.2f TF/s (%5.1f%% of 37.1)\n", name, blocks_per_sm * 8, best,
         flops / best / 1e9, flops / best / 1e9 / 37.1 * 100);
  cudaFree(d);
}

int main() {
  for (int bps : {1, 2, 8}) {
    run<0, 0, 8>("8 DFMA : 0 other", bps);
    run<2, 0, 8>("8 DFMA : 2 IMAD", bps);
    run<4, 0, 8>("8 DFMA : 4 IMAD", bps);
    run<8, 0, 8>("8 DFMA : 8 IMAD", bps);
    run<4, 1, 8>("8 DFMA : 4 FFMA", bps);
    run<8, 1, 8>("8 DFMA : 8 FFMA", bps);
    run<2, 2, 8>("8 DFMA : 2 (LDS.64 + DADD)", bps);
    run<0, 0, 4>("4 DFMA chains (ILP 4) : 0 other", bps);
    run<0, 0, 16>("16 DFMA chains (ILP 16) : 0 other", bps);
  }
  return 0;
}
