// families.cuh — built-in device log-posterior families (the replacement of the user
// closure `fun`).
//   F1 Gaussian LM   README.md:128-139, 356-360; vignettes/advanced-features.Rmd:46-53
//   F2 logistic      vignettes/workflow-with-fmcmc.Rmd:35-41
//   F3 hier. normal  playground/hierarchical-bayes.Rmd:45-51
// Each family splits into a per-observation term (summed over n by many threads)
// and a per-chain "finish" that turns the reduced sum(s) into f(theta).
#pragma once
#include "common.cuh"
#include "softplus.h"

// log(1 + exp(-a)), a >= 0: branch-free FP64 implementation (softplus.h), <= ~2 ulp
__device__ __forceinline__ double softplus_neg(double a) { return fm_softplus_neg(a); }

// min(z, 0) with NaN passed through, as one integer test on the high word: true for every negative
// number (sign bit set, incl. -0 and -inf) and for the default quiet NaNs of either sign.
__device__ __forceinline__ double fm_min0(double z) {
  return ((uint32_t)fm_hi_word(z) > 0x7FF00000u) ? z : 0.0;
}

// Logistic term for one observation: y==1 -> logp, y==0 -> logq, else 0
// (sum(logp[y == 1]) + sum(logq[y == 0]), workflow-with-fmcmc.Rmd:37-39).  With z = eta (y == 1) or
// -eta (y == 0): eta<0 ? eta - log1p(exp(eta)) : -log1p(exp(-eta))  ==  min(z, 0) - log1p(exp(-|z|)).
__device__ __forceinline__ double logistic_term(double eta, double y) {
  const double t = softplus_neg(fabs(eta));
  const double z = (y == 1.0) ? eta : -eta;
  const double v = fm_min0(z) - t;
  return (y == 1.0 || y == 0.0) ? v : 0.0;
}

// Same term when y is known to be exactly +0.0 or 1.0 (ModelParams::y_binary).  Sign flip, min and the
// y test are integer operations on the high word (ALU pipe); the FP64 pipe only runs the softplus and
// one subtraction.
__device__ __forceinline__ double logistic_term_binary(double eta, double y) {
  const long long flip = (fm_hi_word(y) == 0) ? (long long)0x8000000000000000ULL : 0LL;  // y == 0 -> z = -eta
  const double z = fm_bits_to_double(fm_double_to_bits(eta) ^ flip);
  const double t = softplus_neg(fabs(eta));
  return fm_min0(z) - t;
}

// Table-driven variant for the observation-tiled kernel (softplus.h, fm_softplus_tab_core): `tab` is
// the CTA's shared-memory copy of the (S_k, G_k) table, 16 B per entry -> one LDS.128.
template <bool YBIN>
__device__ __forceinline__ double logistic_term_tab(double eta, double y, const double2* __restrict__ tab) {
  double a = fabs(eta);
  a = (fm_hi_word(a) >= 0x40500000) ? (double)FM_SP_AMAX : a;  // >= 64, inf, NaN -> 64 (NaN re-enters via fm_min0)
  const double MAGIC = 6755399441055744.0;                       // 1.5 * 2^52: low word of t = round(32 a)
  const double t = fma(a, (double)FM_SP_H, MAGIC);
  const int k = (int)(uint32_t)fm_double_to_bits(t);
  const double d = fma(t - MAGIC, -1.0 / FM_SP_H, a);            // exact
  const double2 sg = tab[k];
  const double g = fm_softplus_tab_core(d, sg.x, sg.y);
  if (YBIN) {
    const long long flip = (fm_hi_word(y) == 0) ? (long long)0x8000000000000000ULL : 0LL;
    const double z = fm_bits_to_double(fm_double_to_bits(eta) ^ flip);
    return fm_min0(z) - g;
  }
  const double z = (y == 1.0) ? eta : -eta;
  const double v = fm_min0(z) - g;
  return (y == 1.0 || y == 0.0) ? v : 0.0;
}

// sum_i dnorm(r_i, 0, sd, log=TRUE) from ss = sum r_i^2 with R's dnorm edge cases
// (nmath/dnorm.c): sd<0 -> NaN, sd==0 -> +-Inf, !finite(sd) -> -Inf.
__device__ __forceinline__ double gauss_sum_from_ss(double ss, double n, double sd) {
  if (isnan(sd) || isnan(ss)) return NAN;
  if (sd < 0.0) return NAN;
  if (!isfinite(sd)) return -INFINITY;
  if (sd == 0.0) return (ss == 0.0) ? INFINITY : -INFINITY;
  if (!isfinite(ss)) return -INFINITY;
  const double inv = 1.0 / sd;
  return -(n * (FM_LN_SQRT_2PI + log(sd)) + 0.5 * ss * inv * inv);
}

__device__ __forceinline__ double dnorm_log1(double x, double mu, double sd) {
  if (isnan(x) || isnan(mu) || isnan(sd)) return x + mu + sd;
  if (sd < 0.0) return NAN;
  if (!isfinite(sd)) return -INFINITY;
  if (!isfinite(x) && mu == x) return NAN;
  if (sd == 0.0) return (x == mu) ? INFINITY : -INFINITY;
  double z = (x - mu) / sd;
  if (!isfinite(z)) return -INFINITY;
  return -(FM_LN_SQRT_2PI + 0.5 * z * z + log(sd));
}

// Per-chain finish: `s` is the reduced per-observation sum (sum of squares for the
// Gaussian families, the log-likelihood itself for the logistic).
__device__ __forceinline__ double family_finish(const ModelParams& mp, const double* th, double s) {
  switch (mp.family) {
    case FMCMC_FAMILY_GAUSSIAN_LM: {
      double v = gauss_sum_from_ss(s, (double)mp.n_total, th[mp.k - 1]);
      if ((mp.flags & FMCMC_MODEL_GUARD) && !isfinite(v)) v = -INFINITY;  // README.md:135-136
      return v;
    }
    case FMCMC_FAMILY_LOGISTIC: {
      double b2 = 0.0;
      for (int j = 0; j < mp.k; j++) b2 = fma(th[j], th[j], b2);
      if (isnan(b2)) return NAN;  // a NaN parameter makes x %*% beta NaN for every observation
      return s - b2 / (2.0 * mp.h0 * mp.h0);  // - sum(beta^2)/8 for prior sd 2
    }
    case FMCMC_FAMILY_HIER_NORMAL: {
      const int G = mp.n_groups;
      const double gamma = th[G];
      double sigma = 1.0, tau = 1.0;
      if (mp.flags & FMCMC_MODEL_SCALES) { sigma = th[G + 1]; tau = th[G + 2]; }
      double v = gauss_sum_from_ss(s, (double)mp.n_total, sigma);
      double pr = 0.0;
      for (int g = 0; g < G; g++) pr += dnorm_log1(th[g], gamma, tau);
      double du;
      if (isnan(gamma)) du = NAN;
      else du = (mp.h0 <= gamma && gamma <= mp.h1) ? -log(mp.h1 - mp.h0) : -INFINITY;
      return v + pr + du;
    }
  }
  return NAN;
}

// Partial per-observation sum over i = first, first+stride, ... < n for one parameter
// vector `th` (shared or global).  X/y/group may live in shared memory (ld = mp.ld).
__device__ __forceinline__ double family_partial(const ModelParams& mp, const double* __restrict__ X,
                                                 const double* __restrict__ y, const int* __restrict__ grp,
                                                 const double* __restrict__ th, long long first, long long stride) {
  const long long n = mp.n, ld = mp.ld;
  double acc = 0.0;
  switch (mp.family) {
    case FMCMC_FAMILY_GAUSSIAN_LM: {
      const int icpt = (mp.flags & FMCMC_MODEL_INTERCEPT) ? 1 : 0;
      const double b0 = icpt ? th[0] : 0.0;
      double acc2 = 0.0;
      long long i = first;
      for (; i + stride < n; i += 2 * stride) {
        double m0 = b0, m1 = b0;
        for (int j = 0; j < mp.p_x; j++) {
          const double b = th[icpt + j];
          m0 = fma(X[i + j * ld], b, m0);
          m1 = fma(X[i + stride + j * ld], b, m1);
        }
        const double r0 = y[i] - m0, r1 = y[i + stride] - m1;
        acc = fma(r0, r0, acc);
        acc2 = fma(r1, r1, acc2);
      }
      if (i < n) {
        double m0 = b0;
        for (int j = 0; j < mp.p_x; j++) m0 = fma(X[i + j * ld], th[icpt + j], m0);
        const double r0 = y[i] - m0;
        acc = fma(r0, r0, acc);
      }
      return acc + acc2;
    }
    case FMCMC_FAMILY_LOGISTIC: {
      for (long long i = first; i < n; i += stride) {
        double eta = 0.0;
        for (int j = 0; j < mp.p_x; j++) eta = fma(X[i + j * ld], th[j], eta);
        acc += logistic_term(eta, y[i]);
      }
      return acc;
    }
    case FMCMC_FAMILY_HIER_NORMAL: {
      for (long long i = first; i < n; i += stride) {
        const double r = y[i] - th[grp[i]];
        acc = fma(r, r, acc);
      }
      return acc;
    }
  }
  return NAN;
}
