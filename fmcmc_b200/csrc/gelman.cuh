// gelman.cuh — device side of convergence_gelman (R/convergence.R:191-246), i.e. the
// data-parallel part of coda::gelman.diag (third-party; formulae in SURVEY App. A.6):
//   gelman_chain_stats   per chain: mean, variance and the kf x kf covariance over the
//                        window rows; the covariances are summed per CTA in a fixed
//                        chain order (-> W after the cross-GPU all-reduce)
//   gelman_moments       per parameter: muhat and the cross-chain moments the psrf needs
//   gelman_between       B = N * var(t(xbar))  as per-CTA partial sums
// The O(kf^3) scalar finish (chol(W), W^-1 B, top eigenvalue) runs on the host.
#pragma once
#include "common.cuh"

// store: [rows][C][k].  Block b handles chains b, b+gridDim.x, ...
__global__ void gelman_chain_stats_kernel(const double* __restrict__ store, int C, int k, long long row_begin,
                                          long long row_end, const int* __restrict__ fidx, int kf,
                                          double* __restrict__ xbar, double* __restrict__ s2,
                                          double* __restrict__ wpart) {
  extern __shared__ double sh_mean[];  // [kf]
  const long long N = row_end - row_begin;
  const size_t rowlen = (size_t)C * k;
  double* wp = wpart + (size_t)blockIdx.x * kf * kf;
  for (int e = threadIdx.x; e < kf * kf; e += blockDim.x) wp[e] = 0.0;
  for (int c = blockIdx.x; c < C; c += gridDim.x) {
    const double* base = store + (size_t)row_begin * rowlen + (size_t)c * k;
    __syncthreads();
    for (int a = threadIdx.x; a < kf; a += blockDim.x) {
      const int ja = fidx[a];
      double s = 0.0;
      for (long long t = 0; t < N; t++) s += base[t * rowlen + ja];
      double mean = s / (double)N;
      double r = 0.0;  // R's mean(): one refinement pass
      for (long long t = 0; t < N; t++) r += base[t * rowlen + ja] - mean;
      mean += r / (double)N;
      sh_mean[a] = mean;
      xbar[(size_t)c * kf + a] = mean;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kf * kf; e += blockDim.x) {
      const int a = e % kf, b = e / kf;
      const int ja = fidx[a], jb = fidx[b];
      const double ma = sh_mean[a], mb = sh_mean[b];
      double s = 0.0;
      for (long long t = 0; t < N; t++) s = fma(base[t * rowlen + ja] - ma, base[t * rowlen + jb] - mb, s);
      const double v = s / (double)(N - 1);
      wp[e] += v;
      if (a == b) s2[(size_t)c * kf + a] = v;
    }
  }
}

// out[e] = sum_b part[b][e] in block order (deterministic)
__global__ void gelman_wsum_kernel(const double* __restrict__ part, int nblocks, int len, double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= len) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; b++) s += part[(size_t)b * len + e];
  out[e] = s;
}

__device__ __forceinline__ double block_sum_256(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
  return s;
}

// One CTA per free parameter a.  mom[a*8 + ...] = {muhat, mean(s2), var(s2), cov(s2, xbar^2), cov(s2, xbar)}
__global__ void gelman_moments_kernel(const double* __restrict__ xbar, const double* __restrict__ s2, long long m,
                                      int kf, double* __restrict__ mom) {
  __shared__ double red[32];
  const int a = blockIdx.x;
  double sx = 0.0, ss = 0.0, sx2 = 0.0;
  for (long long j = threadIdx.x; j < m; j += blockDim.x) {
    const double x = xbar[j * kf + a];
    sx += x; ss += s2[j * kf + a]; sx2 += x * x;
  }
  const double mx = block_sum_256(sx, red) / (double)m;
  const double ms = block_sum_256(ss, red) / (double)m;
  const double mx2 = block_sum_256(sx2, red) / (double)m;
  double vs = 0.0, c1 = 0.0, c2 = 0.0;
  for (long long j = threadIdx.x; j < m; j += blockDim.x) {
    const double x = xbar[j * kf + a];
    const double d = s2[j * kf + a] - ms;
    vs = fma(d, d, vs);
    c1 = fma(d, x * x - mx2, c1);
    c2 = fma(d, x - mx, c2);
  }
  vs = block_sum_256(vs, red) / (double)(m - 1);
  c1 = block_sum_256(c1, red) / (double)(m - 1);
  c2 = block_sum_256(c2, red) / (double)(m - 1);
  if (threadIdx.x == 0) {
    double* o = mom + (size_t)a * 8;
    o[0] = mx; o[1] = ms; o[2] = vs; o[3] = c1; o[4] = c2;
  }
}

// bpart[block][e] = sum over the block's chains of (xbar_j[a]-mu[a]) (xbar_j[b]-mu[b])
__global__ void gelman_between_kernel(const double* __restrict__ xbar, long long m, int kf,
                                      const double* __restrict__ mom, double* __restrict__ bpart) {
  double* bp = bpart + (size_t)blockIdx.x * kf * kf;
  const long long per = (m + gridDim.x - 1) / gridDim.x;
  const long long j0 = (long long)blockIdx.x * per, j1 = min(m, j0 + per);
  for (int e = threadIdx.x; e < kf * kf; e += blockDim.x) {
    const int a = e % kf, b = e / kf;
    const double ma = mom[(size_t)a * 8], mb = mom[(size_t)b * 8];
    double s = 0.0;
    for (long long j = j0; j < j1; j++) s = fma(xbar[j * kf + a] - ma, xbar[j * kf + b] - mb, s);
    bp[e] = s;
  }
}
