// i8mma_probe.cu — probe of the tcgen05 kind::i8 machinery the split-integer likelihood kernel relies on
// (tiled_i8.cuh): shared-memory operand descriptors (K-major, no swizzle), the instruction descriptor,
// the TMEM accumulator layout read back by tcgen05.ld, and the A-operand-in-TMEM layout written by
// tcgen05.st.  Every variant is checked against an integer reference on the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o i8mma_probe i8mma_probe.cu && ./i8mma_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define M_ROWS 128
#define N_COLS 32
#define KB 2  // K = 32 * KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 = no swizzle
}

__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return true;
    if (clock64() - t0 > 2000000000LL) return false;
  }
}

// mode bit0: A operand from TMEM (tcgen05.st) instead of shared memory; bit1: swap the roles of LBO / SBO
__global__ void __launch_bounds__(128, 1) probe_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B,
                                                        int32_t* __restrict__ D, int mode, int* status) {
  __shared__ __align__(128) int8_t sA[KB][M_ROWS * 32];
  __shared__ __align__(128) int8_t sB[KB][N_COLS * 32];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // canonical K-major no-swizzle layout: core matrix = 8 rows x 16 bytes, contiguous (128 B);
  // element (r, k) at (r / 8) * 256 + (k / 16) * 128 + (r % 8) * 16 + (k % 16)
  for (int e = tid; e < KB * M_ROWS * 32; e += 128) {
    const int kb = e / (M_ROWS * 32), r = (e / 32) % M_ROWS, k = e % 32;
    sA[kb][(r / 8) * 256 + (k / 16) * 128 + (r % 8) * 16 + (k % 16)] = A[r * (32 * KB) + kb * 32 + k];
  }
  for (int e = tid; e < KB * N_COLS * 32; e += 128) {
    const int kb = e / (N_COLS * 32), r = (e / 32) % N_COLS, k = e % 32;
    sB[kb][(r / 8) * 256 + (k / 16) * 128 + (r % 8) * 16 + (k % 16)] = B[r * (32 * KB) + kb * 32 + k];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t tD = tmem, tA = tmem + 32;  // D: columns [0,32), A (TS mode): columns [32, 32 + 8 * KB)
  if (mode & 1) {  // A -> TMEM: thread = row, 8 words per K block, word j = bytes k = 4j .. 4j+3
    for (int kb = 0; kb < KB; kb++) {
      uint32_t w[8];
      const int r = tid;
      for (int j = 0; j < 8; j++) {
        uint32_t v = 0;
        for (int b = 0; b < 4; b++) v |= (uint32_t)(uint8_t)A[r * (32 * KB) + kb * 32 + 4 * j + b] << (8 * b);
        w[j] = v;
      }
      const uint32_t ta = tA + kb * 8 + ((uint32_t)(warp * 32) << 16);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(ta), "r"(w[0]),
                   "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // instruction descriptor: D = S32, A = B = signed int8, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N_COLS >> 3) << 17) | ((uint32_t)(M_ROWS >> 4) << 24);
  if (tid == 0) {
    const uint32_t lbo = (mode & 2) ? 256 : 128, sbo = (mode & 2) ? 128 : 256;
    for (int kb = 0; kb < KB; kb++) {
      const uint64_t da = make_desc(smem_u32(&sA[kb][0]), lbo, sbo);
      const uint64_t db = make_desc(smem_u32(&sB[kb][0]), lbo, sbo);
      const uint32_t acc = kb > 0 ? 1u : 0u;
      if (mode & 1) {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tD),
            "r"(tA + kb * 8), "l"(db), "r"(idesc), "r"(acc)
            : "memory");
      } else {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tD),
            "l"(da), "l"(db), "r"(idesc), "r"(acc)
            : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                 : "memory");
  }
  const bool ok = mbar_wait_bounded(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok) {
    if (tid == 0) status[0] = 1;  // the MMA never completed
  } else {
    for (int c0 = 0; c0 < N_COLS; c0 += 8) {
      uint32_t v[8];
      const uint32_t ta = tD + c0 + ((uint32_t)(warp * 32) << 16);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(ta));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; j++) D[(warp * 32 + lane) * N_COLS + c0 + j] = (int32_t)v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

int main() {
  const int K = 32 * KB;
  int8_t *hA = (int8_t*)malloc(M_ROWS * K), *hB = (int8_t*)malloc(N_COLS * K);
  srand(7);
  for (int i = 0; i < M_ROWS * K; i++) hA[i] = (int8_t)(rand() % 129 - 64);
  for (int i = 0; i < N_COLS * K; i++) hB[i] = (int8_t)(rand() % 129 - 64);
  int32_t* ref = (int32_t*)malloc(sizeof(int32_t) * M_ROWS * N_COLS);
  for (int m = 0; m < M_ROWS; m++)
    for (int n = 0; n < N_COLS; n++) {
      int32_t s = 0;
      for (int k = 0; k < K; k++) s += (int32_t)hA[m * K + k] * (int32_t)hB[n * K + k];
      ref[m * N_COLS + n] = s;
    }
  int8_t *dA, *dB;
  int32_t* dD;
  int* dS;
  cudaMalloc(&dA, M_ROWS * K);
  cudaMalloc(&dB, N_COLS * K);
  cudaMalloc(&dD, sizeof(int32_t) * M_ROWS * N_COLS);
  cudaMalloc(&dS, sizeof(int));
  cudaMemcpy(dA, hA, M_ROWS * K, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, N_COLS * K, cudaMemcpyHostToDevice);
  int32_t* hD = (int32_t*)malloc(sizeof(int32_t) * M_ROWS * N_COLS);
  int rc = 0;
  for (int mode = 0; mode < 4; mode++) {
    cudaMemset(dD, 0xFF, sizeof(int32_t) * M_ROWS * N_COLS);
    cudaMemset(dS, 0, sizeof(int));
    probe_kernel<<<1, 128>>>(dA, dB, dD, mode, dS);
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0;
    cudaMemcpy(&st, dS, sizeof(int), cudaMemcpyDeviceToHost);
    cudaMemcpy(hD, dD, sizeof(int32_t) * M_ROWS * N_COLS, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < M_ROWS * N_COLS; i++) bad += hD[i] != ref[i];
    printf("mode %d (A from %s, LBO/SBO %s): cuda=%s status=%d mismatches=%d / %d   D[0][0..3]=%d %d %d %d  ref=%d %d %d %d  D[1][0]=%d ref=%d D[9][5]=%d ref=%d\n",
           mode, (mode & 1) ? "TMEM" : "SMEM", (mode & 2) ? "swapped" : "as designed", cudaGetErrorString(e), st, bad,
           M_ROWS * N_COLS, hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3], hD[N_COLS], ref[N_COLS],
           hD[9 * N_COLS + 5], ref[9 * N_COLS + 5]);
    if (e != cudaSuccess) { rc = 2; break; }
  }
  return rc;
}
