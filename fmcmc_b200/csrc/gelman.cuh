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

// ---- the scalable pair (kf <= 144): O(C N kf) moments + a tiled batched SYRK -----------------------------------------
// gelman_chain_stats_kernel above walks the window once per (a, b) pair of every chain: kf^2 strided passes.  At
// BASELINE configs[4] (8 192 chains x 128 parameters per GPU, 500-row window) that is ~1.3e11 strided loads.  Here:
//   gelman_chain_moments_kernel  one CTA per chain, two coalesced passes: refined mean (R's mean(): sum / N, then
//                                the mean of the residuals added back) and variance = (q - r^2 / N) / (N - 1),
//                                q / r the sums of squared / plain residuals about the first-pass mean;
//   gelman_syrk_kernel<NB>       sum_j sum_t (x_jt - xbar_j)(x_jt - xbar_j)' as ONE centred SYRK over all (chain, row)
//                                pairs: a CTA keeps a 16 NB x 16 NB accumulator in registers (NB x NB per thread,
//                                strided so the shared-memory reads are broadcasts / conflict-free), streams 16-row
//                                tiles of its chains through shared memory (next tile prefetched into registers
//                                while the current one is multiplied) and writes one kf x kf partial; partials are
//                                summed in block order (gelman_wsum_kernel) -> deterministic.
// Algorithmic work at configs[4]: 3 passes over the 4.2 GB window (2 ms of HBM) + 6.7e10 FP64 FMAs x 2 (full square,
// no symmetry exploited) = 7.3 ms at the FP64 peak.
#define GS_RT 16
__global__ void __launch_bounds__(256)
gelman_chain_moments_kernel(const double* __restrict__ store, int C, int k, long long row_begin, long long row_end,
                            const int* __restrict__ fidx, int kf, double* __restrict__ xbar, double* __restrict__ s2) {
  extern __shared__ double gm_sh[];  // 2 x [G][KP] partial sums (G * KP = 256)
  int KP = 1;
  while (KP < kf && KP < 256) KP <<= 1;
  const int G = 256 / KP, g = threadIdx.x / KP, a0 = threadIdx.x % KP;
  double* part = gm_sh;
  double* part2 = gm_sh + (size_t)G * KP;
  const long long N = row_end - row_begin;
  const size_t rowlen = (size_t)C * k;
  for (int c = blockIdx.x; c < C; c += gridDim.x) {
    const double* base = store + (size_t)row_begin * rowlen + (size_t)c * k;
    for (int ab = 0; ab < kf; ab += KP) {  // kf <= 256: one trip; uniform so that the barriers are not divergent
      const int a = ab + a0;
      const bool on = a < kf;
      const int ja = on ? fidx[a] : 0;
      double s = 0.0;
      if (on)
        for (long long t = g; t < N; t += G) s += base[t * rowlen + ja];
      __syncthreads();
      part[g * KP + a0] = s;
      __syncthreads();
      double tot = 0.0;
      for (int q = 0; q < G; q++) tot += part[q * KP + a0];
      const double mean0 = tot / (double)N;
      double r = 0.0, qq = 0.0;
      if (on)
        for (long long t = g; t < N; t += G) {
          const double d = base[t * rowlen + ja] - mean0;
          r += d;
          qq = fma(d, d, qq);
        }
      __syncthreads();
      part[g * KP + a0] = r;
      part2[g * KP + a0] = qq;
      __syncthreads();
      if (g == 0 && on) {
        double rt = 0.0, qt = 0.0;
        for (int q = 0; q < G; q++) { rt += part[q * KP + a0]; qt += part2[q * KP + a0]; }
        xbar[(size_t)c * kf + a] = mean0 + rt / (double)N;
        s2[(size_t)c * kf + a] = (qt - rt * rt / (double)N) / (double)(N - 1);
      }
    }
  }
}

template <int NB>
__global__ void __launch_bounds__(256)
gelman_syrk_kernel(const double* __restrict__ store, int C, int k, long long row_begin, long long row_end,
                   const int* __restrict__ fidx, int kf, const double* __restrict__ xbar,
                   double* __restrict__ wpart) {
  constexpr int KFP = 16 * NB;
  constexpr int PER = GS_RT * KFP / 256;  // tile elements per thread (= NB)
  __shared__ double tile[GS_RT][KFP];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long N = row_end - row_begin;
  const size_t rowlen = (size_t)C * k;
  double acc[NB][NB];
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) acc[i][j] = 0.0;
  // this thread's slots of a tile: element e = threadIdx.x + 256 q -> (row e / KFP, column e % KFP)
  int col[PER], jcol[PER];
#pragma unroll
  for (int q = 0; q < PER; q++) {
    col[q] = (threadIdx.x + 256 * q) % KFP;
    jcol[q] = col[q] < kf ? fidx[col[q]] : -1;
  }
  const long long ntile = (N + GS_RT - 1) / GS_RT;
  for (int c = blockIdx.x; c < C; c += gridDim.x) {
    const double* base = store + (size_t)row_begin * rowlen + (size_t)c * k;
    double mean[PER], nxt[PER];
#pragma unroll
    for (int q = 0; q < PER; q++) mean[q] = jcol[q] >= 0 ? xbar[(size_t)c * kf + col[q]] : 0.0;
    auto fetch = [&](long long tl) {
#pragma unroll
      for (int q = 0; q < PER; q++) {
        const long long t = tl * GS_RT + (threadIdx.x + 256 * q) / KFP;
        nxt[q] = (jcol[q] >= 0 && t < N) ? base[t * rowlen + jcol[q]] - mean[q] : 0.0;
      }
    };
    fetch(0);
    for (long long tl = 0; tl < ntile; tl++) {
      __syncthreads();  // previous tile consumed
#pragma unroll
      for (int q = 0; q < PER; q++) tile[(threadIdx.x + 256 * q) / KFP][col[q]] = nxt[q];
      __syncthreads();
      if (tl + 1 < ntile) fetch(tl + 1);  // in flight while this tile is multiplied
#pragma unroll 4
      for (int tt = 0; tt < GS_RT; tt++) {
        double av[NB], bv[NB];
#pragma unroll
        for (int i = 0; i < NB; i++) { av[i] = tile[tt][ty + 16 * i]; bv[i] = tile[tt][tx + 16 * i]; }
#pragma unroll
        for (int i = 0; i < NB; i++)
#pragma unroll
          for (int j = 0; j < NB; j++) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
      }
    }
  }
  double* wp = wpart + (size_t)blockIdx.x * kf * kf;
  const double inv = 1.0 / (double)(N - 1);
#pragma unroll
  for (int i = 0; i < NB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      const int a = ty + 16 * i, b = tx + 16 * j;
      if (a < kf && b < kf) wp[a + (size_t)b * kf] = acc[i][j] * inv;
    }
}

// out[e] = sum_b part[b][e] in a fixed order (deterministic): eight interleaved partial sums, so that the loads of a trip are
// independent and the add chain is nblocks / 8 long (one serial chain over ~1 200 blocks took 25 us per launch)
__global__ void gelman_wsum_kernel(const double* __restrict__ part, int nblocks, int len, double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= len) return;
  double s[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  int b = 0;
  for (; b + 8 <= nblocks; b += 8) {
#pragma unroll
    for (int q = 0; q < 8; q++) s[q] += part[(size_t)(b + q) * len + e];
  }
  for (int q = 0; b < nblocks; b++, q++) s[q] += part[(size_t)b * len + e];
  out[e] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
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

// rm_invariant (R/convergence.R:169-186, quirk D9) tests ONE number: stats::sd() of every element of rbind(chains) - all
// rows accumulated so far, all (free) parameters, all chains - squared, against 1e-10.  One streaming pass over the store:
// every block sums d = x - shift and d^2 with the same shift (the store's first element) over a grid-strided range;
// part[block] = (count, sum d, sum d^2); the host adds the blocks in order.
__global__ void __launch_bounds__(256)
gelman_pooled_kernel(const double* __restrict__ store, int C, int k, long long rows, const int* __restrict__ fidx, int kf,
                     double* __restrict__ part) {
  __shared__ double red[32];
  const double shift = store[fidx[0]];
  const long long per_row = (long long)C * kf, total = rows * per_row;
  double sd = 0.0, sdd = 0.0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long t = e / per_row, r = e - t * per_row;
    const int c = (int)(r / kf), a = (int)(r - (long long)c * kf);
    const double d = store[((size_t)t * C + c) * k + fidx[a]] - shift;
    sd += d;
    sdd = fma(d, d, sdd);
  }
  const double bs = block_sum_256(sd, red), bq = block_sum_256(sdd, red);
  if (threadIdx.x == 0) {
    const long long first = blockIdx.x * (long long)blockDim.x;
    long long cnt = 0;
    if (first < total) {  // elements this block visited: threads first .. first + 255 of every grid stride
      const long long stride = (long long)gridDim.x * blockDim.x;
      const long long full = (total - first) / stride, rem = (total - first) - full * stride;
      cnt = full * blockDim.x + (rem < (long long)blockDim.x ? rem : (long long)blockDim.x);
    }
    part[blockIdx.x * 3 + 0] = (double)cnt;
    part[blockIdx.x * 3 + 1] = bs;
    part[blockIdx.x * 3 + 2] = bq;
  }
}

// ---- effective sample size on the device (the reference itself only prints coda's time-series SE) ----
// One CTA per (chain, free parameter) series x_1..x_N taken from the sample store: centred series in shared memory,
// autocovariances gamma(l) = (1/N) sum_t x_t x_{t+l} for l = 0 .. L (thread <-> lag, every product read from shared
// memory), then Geyer's initial positive sequence on the pairs Gamma_m = rho(2m) + rho(2m + 1): tau = -1 + 2 sum_m Gamma_m
// up to the first non-positive pair, ESS = N / tau.  O(N L) FP64 multiply-adds per series; L = min(N - 1, max_lag).
// If the pairs are still positive at the last lag the estimate is truncated there (reported through `truncated`).
__global__ void __launch_bounds__(256)
store_ess_kernel(const double* __restrict__ store, int C, int k, long long row_begin, long long row_end,
                 const int* __restrict__ fidx, int kf, int max_lag, double* __restrict__ ess, int* __restrict__ truncated) {
  extern __shared__ double es_sh[];  // [N] centred series, then [L + 2] autocovariances
  __shared__ double red[32];
  const long long N = row_end - row_begin;
  const int c = blockIdx.x / kf, a = blockIdx.x % kf;
  const int L = (int)min((long long)max_lag, N - 1);
  double* x = es_sh;
  double* g = es_sh + N;
  const size_t rowlen = (size_t)C * k;
  const double* base = store + (size_t)row_begin * rowlen + (size_t)c * k + fidx[a];
  double s = 0.0;
  for (long long t = threadIdx.x; t < N; t += blockDim.x) { const double v = base[t * rowlen]; x[t] = v; s += v; }
  const double mean = block_sum_256(s, red) / (double)N;
  __syncthreads();
  for (long long t = threadIdx.x; t < N; t += blockDim.x) x[t] -= mean;
  __syncthreads();
  for (int l = threadIdx.x; l <= L + 1; l += blockDim.x) {
    double acc = 0.0;
    if (l <= L)
      for (long long t = 0; t + l < N; t++) acc = fma(x[t], x[t + l], acc);
    g[l] = acc / (double)N;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double g0 = g[0];
    double tau = -1.0;
    int trunc = 0;
    if (g0 > 0.0) {
      int m = 0;
      for (;; m++) {
        if (2 * m + 1 > L) { trunc = 1; break; }
        const double pair = (g[2 * m] + g[2 * m + 1]) / g0;
        if (!(pair > 0.0)) break;
        tau += 2.0 * pair;
      }
      if (tau < 1.0 / (double)N) tau = 1.0 / (double)N;
      ess[(size_t)c * kf + a] = (double)N / tau;
    } else {
      ess[(size_t)c * kf + a] = 0.0;  // a constant series carries no information
    }
    if (trunc) atomicOr(truncated, 1);
  }
}
