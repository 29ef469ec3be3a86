// i8_epilogue_rate.cu — how fast can the SM run path 4's per-evaluation instruction stream (integer merge + I2F +
// 14 FP64 softplus).  NOTE: synthesising the six accumulator words costs ~6 integer instructions (~12 clk) per
// evaluation here that the kernel does not pay (its words come from tcgen05.ld). when nothing else is in the way (no TMEM, no MMA, no barriers)?  W warps per SM, CH evaluations
// interleaved per warp, inputs synthesised in registers.  Reports clk per warp-evaluation per scheduler; the FP64
// pipe alone needs 38 (19 instructions x 2 clk).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../fmcmc_b200/csrc -o i8_epilogue_rate i8_epilogue_rate.cu
#include <cstdio>
#include "tiled_i8.cuh"

// reassembly variants (all exact up to the 2 dropped bits of the int64 form)
__device__ __forceinline__ double cvt32(int v) { return __hiloint2double(0x43300000, v ^ (int)0x80000000) - (4503599627370496.0 + 2147483648.0); }
template <int CH>
__device__ __forceinline__ double merge_variant(const uint32_t (&a)[6][CH], int e, int variant) {
  const int v0 = (int)a[0][e] * 128 + (int)a[1][e], v1 = (int)a[2][e] * 128 + (int)a[3][e], v2 = (int)a[4][e] * 128 + (int)a[5][e];
  if (variant == 1) return fma(fma(cvt32(v0), 16384.0, cvt32(v1)), 16384.0, cvt32(v2));                 // 3 magic32 + 2 DFMA
  if (variant == 2) return fma(fma(__int2double_rn(v0), 16384.0, __int2double_rn(v1)), 16384.0, __int2double_rn(v2));  // 3 I2F.S32 + 2 DFMA
  if (variant == 3) return __ll2double_rn((long long)v0 * (1 << 28) + (long long)v1 * 16384 + v2);     // I2F.F64.S64
  if (variant == 4) {  // all six words converted separately: 6 LOP3 + 6 DADD + 5 DFMA, no IMAD
    double t = cvt32((int)a[0][e]);
#pragma unroll
    for (int d = 1; d < 6; d++) t = fma(t, 128.0, cvt32((int)a[d][e]));
    return t;
  }
  if (variant == 5) {  // FP32 route: a_d < 2^24 is exact in float; pairs merged in FP64
    const double d0 = (double)(float)(int)a[0][e], d1 = (double)(float)(int)a[1][e];
    return fma(d0, 128.0, d1);
  }
  return 0.0;
}

template <int CH, int MODE>
__global__ void __launch_bounds__(1024, 1) k(double* out, const double* gtab, int iters, long long* cyc) {
  extern __shared__ double2 tab[];   // FM_SP8_ENTRIES (fine table)
  for (int e = threadIdx.x; e < FM_SP8_ENTRIES; e += blockDim.x) tab[e] = reinterpret_cast<const double2*>(gtab)[e];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const double csc = 1.0 / 35184372088832.0;  // 2^-45: eta = t * csc of order 1
  double acc = 0.0, acc2 = 0.0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    uint32_t a[6][CH];
#pragma unroll
    for (int d = 0; d < 6; d++)
#pragma unroll
      for (int e = 0; e < CH; e++) a[d][e] = (uint32_t)((lane * 2654435761u + it * 40503u + d * 977u + e * 131u) >> 13) - 262144u;
#pragma unroll
    for (int d = 0; d < 6; d++)
#pragma unroll
      for (int e = 0; e < CH; e++) asm volatile("" : "+r"(a[d][e]));
#pragma unroll
    for (int e = 0; e < CH; e++) {
      if (MODE == 0) {            // the real thing
        const double t = i8_assemble<6, CH>(a, e);
        i8_logistic_even_t<2, false>(t, csc, 0x40440000 + (45 << 20), acc, acc2, tab);
      } else if (MODE == 1) {     // FP64 part only: the argument comes from a cheap int -> double trick
        const double t = __hiloint2double(0x43300000, (int)a[0][e]) - 4503599627370496.0;
        i8_logistic_even_t<2, false>(t, 9.5367431640625e-07, 0x40440000 + (20 << 20), acc, acc2, tab);
      } else if (MODE == 2) {     // integer merge + conversion only (the kernel's form)
        acc += i8_assemble<6, CH>(a, e);
      } else {
        acc += merge_variant<CH>(a, e, MODE - 10);
      }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + acc2;
  if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double *out, *gtab;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 8);
  double* htab = new double[2 * FM_SP8_ENTRIES];
  fm_softplus_table8_fill(htab);
  cudaMalloc(&gtab, 2 * FM_SP8_ENTRIES * 8);
  cudaMemcpy(gtab, htab, 2 * FM_SP8_ENTRIES * 8, cudaMemcpyHostToDevice);
  const int iters = 2000;
  auto run = [&](auto kern, int warps, int CH, const char* name) {
    long long hc = 0;
    for (int rep = 0; rep < 2; rep++) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FM_SP8_ENTRIES * 16);
      kern<<<148, warps * 32, FM_SP8_ENTRIES * 16>>>(out, gtab, iters, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
      cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
    }
    const double wevals_per_sched = (double)iters * CH * warps / 4.0;
    printf("%-44s %2d warps/SM, %d in flight: %7.1f clk per warp-evaluation per scheduler\n", name, warps, CH, hc / wevals_per_sched);
  };
  for (int warps : {8, 16}) {
    run(k<8, 0>, warps, 8, "merge + I2F + softplus (the epilogue)");
    run(k<4, 0>, warps, 4, "merge + I2F + softplus (the epilogue)");
    run(k<8, 1>, warps, 8, "softplus only (14 FP64 + LDS + 3 int)");
    run(k<4, 1>, warps, 4, "softplus only (14 FP64 + LDS + 3 int)");
    run(k<8, 2>, warps, 8, "merge only: int32 pairs, int64, I2F.S64 (kernel)");
    if (warps == 16) {
      run(k<8, 11>, warps, 8, "merge only: int32 pairs, 3 magic32, 2 DFMA");
      run(k<8, 12>, warps, 8, "merge only: int32 pairs, 3 I2F.S32, 2 DFMA");
      run(k<8, 14>, warps, 8, "merge only: 6 magic32, 5 DFMA");
      run(k<8, 15>, warps, 8, "2 I2F.F32 + 2 F2F + 1 DFMA (partial, cost probe)");
    }
  }
  return 0;
}
