/*
 * oracle/fmcmc_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * A CPU restatement, in plain C, of the hot path of USCbiostats/fmcmc v0.6-0
 * (100% interpreted R; R is not installed in this image, so the reference
 * itself cannot be executed).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (fmcmc_b200/) never does.
 *
 * Parity pinning (see tests/test_oracle_readme_golden.py):
 *   - README.md:183-201  posterior summary of the seed-1215 run,
 *   - README.md:315-339, 388-412  the two 13-value Gelman-Rubin R traces,
 *   are regenerated bit-for-bit-in-print through oracle/r_rng.c (R's RNG) +
 *   this file (loop, kernel_normal, kernel_normal_reflective, bulk loop,
 *   coda::gelman.diag);
 *   - inst/tinytest/test-kernel_adapt.R:33-55 (cov_recursive == cov KAT).
 *   kernel_adapt after warm-up (MASS::mvrnorm's LAPACK eigenvectors),
 *   kernel_ram's nearPD failure path and rt()/rgamma streams are third-party
 *   and UNPINNED; they are restated from their published
 *   algorithms.
 *
 * Every function cites the reference file:line it follows.  Arithmetic order
 * follows the R expressions literally and the file is compiled with
 * -ffp-contract=off; sums use long double like R's rsum().
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "../include/fmcmc_b200.h"

#define M_LN_SQRT_2PI_ 0.918938533204672741780329736406 /* nmath: log(sqrt(2*pi)) */

static void set_err(char* err, size_t errlen, const char* fmt, ...)
    __attribute__((format(printf, 3, 4)));
#include <stdarg.h>
static void set_err(char* err, size_t errlen, const char* fmt, ...) {
  if (!err || !errlen) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err, errlen, fmt, ap);
  va_end(ap);
}

/* ======================================================================== */
/* base R pieces                                                             */
/* ======================================================================== */

/* nmath/dnorm.c dnorm4(x, mu, sigma, give_log = TRUE)  [base R, restated] */
static double r_dnorm_log(double x, double mu, double sigma) {
  if (isnan(x) || isnan(mu) || isnan(sigma)) return x + mu + sigma;
  if (sigma < 0) return NAN;
  if (!isfinite(sigma)) return -INFINITY;
  if (!isfinite(x) && mu == x) return NAN;
  if (sigma == 0) return (x == mu) ? INFINITY : -INFINITY;
  x = (x - mu) / sigma;
  if (!isfinite(x)) return -INFINITY;
  x = fabs(x);
  if (x >= 2 * sqrt(DBL_MAX)) return -INFINITY;
  return -(M_LN_SQRT_2PI_ + 0.5 * x * x + log(sigma));
}

/* nmath/dunif.c, log = TRUE */
static double r_dunif_log(double x, double a, double b) {
  if (isnan(x) || isnan(a) || isnan(b)) return x + a + b;
  if (b <= a) return NAN;
  if (a <= x && x <= b) return -log(b - a);
  return -INFINITY;
}

/* arithmetic.c myfmod / myfloor: R's %% and %/% on doubles */
static double r_fmod(double x1, double x2) {
  if (x2 == 0.0) return NAN;
  if (fabs(x2) * DBL_EPSILON > 1 && isfinite(x1) && fabs(x1) <= fabs(x2)) {
    return (fabs(x1) == fabs(x2)) ? 0
           : ((x1 < 0 && x2 > 0) || (x2 < 0 && x1 > 0)) ? x1 + x2 : x1;
  }
  double q = x1 / x2;
  long double tmp = (long double)x1 - floor(q) * (long double)x2;
  return (double)(tmp - floorl(tmp / x2) * x2);
}
static double r_intdiv(double x1, double x2) {
  double q = x1 / x2;
  if (x2 == 0.0 || fabs(q) * DBL_EPSILON > 1 || !isfinite(q)) return q;
  if (fabs(q) < 1)
    return (q < 0) ? -1 : ((x1 < 0 && x2 > 0) || (x1 > 0 && x2 < 0) ? -1 : 0);
  long double tmp = (long double)x1 - floor(q) * (long double)x2;
  return (double)(floor(q) + floorl(tmp / x2));
}

/* ======================================================================== */
/* Log-posterior families (the user closures the reference documents)        */
/* ======================================================================== */

int32_t fmcmc_oracle_nparams(const fmcmc_model_desc* d) {
  switch (d->family) {
    case FMCMC_FAMILY_GAUSSIAN_LM:
      return d->p_x + ((d->flags & FMCMC_MODEL_INTERCEPT) ? 1 : 0) + 1;
    case FMCMC_FAMILY_LOGISTIC:
      return d->p_x;
    case FMCMC_FAMILY_HIER_NORMAL:
      return d->n_groups + 1 + ((d->flags & FMCMC_MODEL_SCALES) ? 2 : 0);
  }
  return -1;
}

/* README.md:128-139 (guarded) / 356-360 (unguarded);
 * vignettes/advanced-features.Rmd:46-53 is the same with three terms. */
static double ll_gaussian_lm(const fmcmc_model_desc* d, const double* th) {
  const int icpt = (d->flags & FMCMC_MODEL_INTERCEPT) ? 1 : 0;
  const int k = d->p_x + icpt + 1;
  const double sd = th[k - 1];
  long double s = 0.0L;
  for (int64_t i = 0; i < d->n; i++) {
    double mu;
    int j0 = 0;
    if (icpt) {
      mu = th[0];
    } else {
      mu = d->X[i] * th[0];
      j0 = 1;
    }
    for (int j = j0; j < d->p_x; j++) mu = mu + d->X[i + (int64_t)j * d->n] * th[icpt + j];
    s += r_dnorm_log(d->y[i] - mu, 0.0, sd);
  }
  double v = (double)s;
  if ((d->flags & FMCMC_MODEL_GUARD) && !isfinite(v)) return -INFINITY;
  return v;
}

/* vignettes/workflow-with-fmcmc.Rmd:35-41 (prior sd 2 => sum(beta^2)/8) */
static double ll_logistic(const fmcmc_model_desc* d, const double* th) {
  const int k = d->p_x;
  long double s1 = 0.0L, s0 = 0.0L;
  for (int64_t i = 0; i < d->n; i++) {
    long double e = 0.0L; /* x %*% beta: BLAS dgemv; accumulate then round */
    double eta = 0.0;
    for (int j = 0; j < k; j++) eta += d->X[i + (int64_t)j * d->n] * th[j];
    (void)e;
    if (d->y[i] == 1.0) {
      s1 += (eta < 0) ? eta - log1p(exp(eta)) : -log1p(exp(-eta));
    } else if (d->y[i] == 0.0) {
      s0 += (eta < 0) ? -log1p(exp(eta)) : -eta - log1p(exp(-eta));
    }
  }
  long double b2 = 0.0L;
  for (int j = 0; j < k; j++) b2 += th[j] * th[j];
  double logl = (double)s1 + (double)s0;
  double psd = d->hyper[0];
  return logl - (double)b2 / (2.0 * psd * psd);
}

/* playground/hierarchical-bayes.Rmd:45-51; with FMCMC_MODEL_SCALES the two
 * unit standard deviations become parameters. */
static double ll_hier_normal(const fmcmc_model_desc* d, const double* th) {
  const int G = d->n_groups;
  const double gamma = th[G];
  double sigma = 1.0, tau = 1.0;
  if (d->flags & FMCMC_MODEL_SCALES) {
    sigma = th[G + 1];
    tau = th[G + 2];
  }
  long double s = 0.0L;
  for (int64_t i = 0; i < d->n; i++) s += r_dnorm_log(d->y[i], th[d->group[i]], sigma);
  long double s2 = 0.0L;
  for (int g = 0; g < G; g++) s2 += r_dnorm_log(th[g], gamma, tau);
  return (double)s + (double)s2 + r_dunif_log(gamma, d->hyper[0], d->hyper[1]);
}

double fmcmc_oracle_logpost(const fmcmc_model_desc* d, const double* th) {
  switch (d->family) {
    case FMCMC_FAMILY_GAUSSIAN_LM: return ll_gaussian_lm(d, th);
    case FMCMC_FAMILY_LOGISTIC: return ll_logistic(d, th);
    case FMCMC_FAMILY_HIER_NORMAL: return ll_hier_normal(d, th);
  }
  return NAN;
}

/* ======================================================================== */
/* R/kernel.R:450-493 reflect_on_boundaries                                  */
/* ======================================================================== */
void fmcmc_oracle_reflect(int k, double* x, const double* lb, const double* ub,
                          const uint8_t* which) {
  for (int j = 0; j < k; j++) {
    if (which && !which[j]) continue;
    double d = ub[j] - lb[j];
    if (x[j] > ub[j]) {
      double d_above = x[j] - ub[j];
      double odd = r_fmod(r_intdiv(d_above, d), 2.0);
      d_above = r_fmod(d_above, d);
      x[j] = (lb[j] + d_above) * odd + (ub[j] - d_above) * (1 - odd);
    } else if (x[j] < lb[j]) {
      double d_below = lb[j] - x[j];
      double odd = r_fmod(r_intdiv(d_below, d), 2.0);
      d_below = r_fmod(d_below, d);
      x[j] = (ub[j] - d_below) * odd + (lb[j] + d_below) * (1 - odd);
    }
  }
}

/* ======================================================================== */
/* R/recursive.R:124-139 mean_recursive, 63-120 cov_recursive (vector form)  */
/* ======================================================================== */
static void mean_rec1(int k, const double* x, const double* mprev, double t, double* m) {
  for (int a = 0; a < k; a++) m[a] = (mprev[a] * t + x[a]) / (t + 1);
}
/* cov (k x k col-major) updated in place; Ik is the matrix passed as `Ik` */
static void cov_rec1(int k, const double* x, double* cov, const double* m, const double* mprev,
                     double t, double eps, double Sd, const double* Ik) {
  for (int b = 0; b < k; b++)
    for (int a = 0; a < k; a++) {
      double inner = t * (mprev[a] * mprev[b]) - (t + 1) * (m[a] * m[b]) + x[a] * x[b] +
                     eps * Ik[a + b * k];
      cov[a + b * k] = (t - 1) / t * cov[a + b * k] + Sd / t * inner;
    }
}

/* matrix mode (R/recursive.R:78-110, 128-137): rows chained with t + i - 1 */
void fmcmc_oracle_cov_recursive(int k, int64_t rows, const double* X /*[rows][k]*/,
                                const double* mean_prev, const double* cov_prev, double t,
                                double eps, double Sd, const double* Ik, double* mean_out,
                                double* cov_out) {
  double* m = (double*)malloc(sizeof(double) * k);
  double* mp = (double*)malloc(sizeof(double) * k);
  double* eye = NULL;
  if (!Ik) {
    eye = (double*)calloc((size_t)k * k, sizeof(double));
    for (int a = 0; a < k; a++) eye[a + a * k] = 1.0;
    Ik = eye;
  }
  memcpy(mp, mean_prev, sizeof(double) * k);
  memcpy(cov_out, cov_prev, sizeof(double) * k * k);
  for (int64_t i = 0; i < rows; i++) {
    double ti = t + (double)i;
    mean_rec1(k, X + i * k, mp, ti, m);
    cov_rec1(k, X + i * k, cov_out, m, mp, ti, eps, Sd, Ik);
    memcpy(mp, m, sizeof(double) * k);
  }
  memcpy(mean_out, mp, sizeof(double) * k);
  free(m);
  free(mp);
  free(eye);
}

/* ======================================================================== */
/* small dense linear algebra                                                */
/* ======================================================================== */

/* lower Cholesky of a symmetric k x k (col-major); returns 0 ok, j+1 if pivot j <= 0
 * (LAPACK dpotrf semantics used by R's chol()). */
static int chol_lower(int k, const double* A, double* L) {
  memset(L, 0, sizeof(double) * k * k);
  for (int j = 0; j < k; j++) {
    double s = A[j + j * k];
    for (int p = 0; p < j; p++) s -= L[j + p * k] * L[j + p * k];
    if (!(s > 0.0)) return j + 1;
    double ljj = sqrt(s);
    L[j + j * k] = ljj;
    for (int i = j + 1; i < k; i++) {
      double v = A[i + j * k];
      for (int p = 0; p < j; p++) v -= L[i + p * k] * L[j + p * k];
      L[i + j * k] = v / ljj;
    }
  }
  return 0;
}

/* eigen(Sigma, symmetric = TRUE) as MASS::mvrnorm consumes it (R/kernel_adapt.R:173-178).
 * Third-party: MASS (no pin, DESCRIPTION) -> base R eigen() -> LAPACK dsyevr.  What is restated:
 *   - eigenvalues DEcreasing, obtained as R does by REVERSING LAPACK's ascending order, so that ties
 *     come out in descending index order (eps*I, the whole warm-up, gives the exchange matrix, i.e.
 *     the draw is sqrt(eps) * (z_k, ..., z_1): pinned by the logpost trace the reference publishes in
 *     man/figures/get_-1.png, tests/test_oracle_readme_golden.py);
 *   - the SIGN of each eigenvector is an artefact of LAPACK's MRRR internals (dstemr: the component at
 *     the twist index, in the tridiagonal basis, is positive) and, for the exactly repeated eigenvalue
 *     of the first adapted Sigma (eps*I + rank one), so is the basis itself.  It is not reproducible
 *     without the R installation's own LAPACK binary (README.md:268-269 is not reproduced by OpenBLAS'
 *     dsyevr either).  Convention here: the largest |component| of every
 *     eigenvector is positive (lowest index on ties).
 * Algorithm: cyclic-by-row Jacobi, rotations skipped element-wise when |a_pq| <= 1e-17 sqrt|a_pp a_qq|,
 * sweeps until one applies no rotation; only + - * / sqrt in a fixed order, so that the CUDA head
 * (propose.cuh eigen_factor_warp) reproduces it bit for bit.  A is destroyed. */
static void jacobi_eigen(int k, double* A, double* ev, double* V) {
  for (int i = 0; i < k * k; i++) V[i] = 0.0;
  for (int i = 0; i < k; i++) V[i + i * k] = 1.0;
  for (int sweep = 0; sweep < 64; sweep++) {
    int rotated = 0;
    for (int p = 0; p < k - 1; p++)
      for (int q = p + 1; q < k; q++) {
        double apq = A[p + q * k];
        double app = A[p + p * k], aqq = A[q + q * k];
        if (apq == 0.0 || fabs(apq) <= 1e-17 * sqrt(fabs(app * aqq))) continue;
        rotated = 1;
        double theta = (aqq - app) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int r = 0; r < k; r++) { /* columns p,q */
          double arp = A[r + p * k], arq = A[r + q * k];
          A[r + p * k] = c * arp - s * arq;
          A[r + q * k] = s * arp + c * arq;
        }
        for (int r = 0; r < k; r++) { /* rows p,q */
          double apr = A[p + r * k], aqr = A[q + r * k];
          A[p + r * k] = c * apr - s * aqr;
          A[q + r * k] = s * apr + c * aqr;
        }
        for (int r = 0; r < k; r++) {
          double vrp = V[r + p * k], vrq = V[r + q * k];
          V[r + p * k] = c * vrp - s * vrq;
          V[r + q * k] = s * vrp + c * vrq;
        }
      }
    if (!rotated) break;
  }
  for (int i = 0; i < k; i++) ev[i] = A[i + i * k];
  /* selection sort, decreasing; among equal eigenvalues the HIGHER original index first */
  int idx[k];
  for (int i = 0; i < k; i++) idx[i] = i;
  for (int i = 0; i < k - 1; i++) {
    int m = i;
    for (int j = i + 1; j < k; j++)
      if (ev[j] > ev[m] || (ev[j] == ev[m] && idx[j] > idx[m])) m = j;
    if (m != i) {
      double tv = ev[i]; ev[i] = ev[m]; ev[m] = tv;
      int ti = idx[i]; idx[i] = idx[m]; idx[m] = ti;
      for (int r = 0; r < k; r++) {
        double tt = V[r + i * k]; V[r + i * k] = V[r + m * k]; V[r + m * k] = tt;
      }
    }
  }
  for (int j = 0; j < k; j++) { /* sign convention */
    int m = 0;
    for (int r = 1; r < k; r++)
      if (fabs(V[r + j * k]) > fabs(V[m + j * k])) m = r;
    if (V[m + j * k] < 0.0)
      for (int r = 0; r < k; r++) V[r + j * k] = -V[r + j * k];
  }
}
/* test hook: eigen-decomposition of a col-major symmetric matrix as the EIGEN draw uses it */
void fmcmc_oracle_eigen(int k, const double* Sigma, double* ev, double* V) {
  double A[k * k];
  memcpy(A, Sigma, sizeof(A));
  jacobi_eigen(k, A, ev, V);
}

/* ======================================================================== */
/* Philox4x32-10 production stream (shared definition with the CUDA kernels)  */
/* ======================================================================== */
static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                 uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline double u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}
/* two U(0,1) for (global chain, run, row, slot) */
void fmcmc_oracle_philox_u2(uint64_t seed, uint32_t chain, uint32_t run, uint32_t row,
                            uint32_t slot, double* u0, double* u1) {
  uint32_t o[4];
  philox4x32_10(chain, run, row, slot, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  *u0 = u53(o[0], o[1]);
  *u1 = u53(o[2], o[3]);
}
extern double r_qnorm(double p); /* AS 241, oracle/r_rng.c */

/* slot map (kept identical in fmcmc_b200/csrc/philox.cuh):
 *   slot 0            : u0 -> accept uniform, u1 -> random-scheme coordinate
 *   slot 1 + j/2      : standard normal / uniform for the j-th active coordinate
 *   RAM (rt(k,k))     : slot 1 + j*32 + a, a in [0,31): Marsaglia-Tsang attempt a
 *                       (u0 -> normal, u1 -> uniform); slot 1 + j*32 + 31: u0 ->
 *                       numerator normal, u1 -> boost uniform (shape < 1)
 */
#define PLAN_RUN 0xFFFFFFFFu

typedef struct {
  int mode;
  const double* zrow; /* FED: this row's slots */
  uint64_t seed;
  uint32_t chain, run, row;
} draw_src;

static double draw_u01(const draw_src* s, int j) {
  if (s->mode == FMCMC_STREAM_FED) return s->zrow[j];
  double u0, u1;
  fmcmc_oracle_philox_u2(s->seed, s->chain, s->run, s->row, 1u + (uint32_t)(j / 2), &u0, &u1);
  return (j & 1) ? u1 : u0;
}
static double draw_z(const draw_src* s, int j) {
  if (s->mode == FMCMC_STREAM_FED) return s->zrow[j];
  return r_qnorm(draw_u01(s, j));
}
/* Student t with df degrees of freedom (kernel_ram's default qfun = rt(k, k)) */
static double draw_t(const draw_src* s, int j, double df) {
  if (s->mode == FMCMC_STREAM_FED) return s->zrow[j];
  double u0, u1;
  uint32_t base = 1u + (uint32_t)j * 32u;
  double a = 0.5 * df, boost = 1.0;
  fmcmc_oracle_philox_u2(s->seed, s->chain, s->run, s->row, base + 31u, &u0, &u1);
  double znum = r_qnorm(u0);
  if (a < 1.0) {
    boost = pow(u1, 1.0 / a);
    a += 1.0;
  }
  double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d), g = d;
  for (uint32_t at = 0; at < 31u; at++) {
    fmcmc_oracle_philox_u2(s->seed, s->chain, s->run, s->row, base + at, &u0, &u1);
    double x = r_qnorm(u0);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    g = d * v;
    if (log(u1) < 0.5 * x * x + d - d * v + d * log(v)) break;
  }
  double chi2 = 2.0 * g * boost;
  return znum / sqrt(chi2 / df);
}

/* ======================================================================== */
/* kernels                                                                    */
/* ======================================================================== */
typedef struct {
  const fmcmc_kernel_spec* ks;
  int k, kf;
  int* free_idx;     /* which(!fixed), 0-based */
  int64_t* istate;   /* this chain */
  double* dstate;    /* this chain */
  const int32_t* seq;/* this chain's planned random sequence (1-based) or NULL */
} kchain;

int64_t fmcmc_oracle_state_len(int32_t type, int32_t k, int32_t kf) {
  switch (type) {
    case FMCMC_KERNEL_ADAPT: return (int64_t)kf * kf + kf;
    case FMCMC_KERNEL_RAM: return (int64_t)kf * kf;
    case FMCMC_KERNEL_NMIRROR:
    case FMCMC_KERNEL_UMIRROR: return 3 * (int64_t)k;
  }
  return 0;
}

/* active coordinates of row i (1-based) -> act[], returns count.
 * R/kernel.R:66-133 plan_update_sequence (joint/ordered/random/explicit). */
static int active_coords(const kchain* kc, const draw_src* ds, int64_t i, int* act) {
  const fmcmc_kernel_spec* ks = kc->ks;
  switch (ks->scheme) {
    case FMCMC_SCHEME_JOINT:
      for (int j = 0; j < kc->kf; j++) act[j] = kc->free_idx[j];
      return kc->kf;
    case FMCMC_SCHEME_ORDERED:
      act[0] = kc->free_idx[(i - 1) % kc->kf];
      return 1;
    case FMCMC_SCHEME_EXPLICIT:
      act[0] = ks->order[(i - 1) % ks->order_len] - 1;
      return 1;
    case FMCMC_SCHEME_RANDOM:
      if (kc->seq) {
        act[0] = kc->seq[i - 1] - 1;
      } else {
        double u0, u1;
        fmcmc_oracle_philox_u2(ds->seed, ds->chain, PLAN_RUN, (uint32_t)i, 0u, &u0, &u1);
        int pick = (int)(u1 * kc->kf);
        if (pick >= kc->kf) pick = kc->kf - 1;
        act[0] = kc->free_idx[pick];
      }
      return 1;
  }
  return 0;
}

typedef struct {
  const double* ans; /* [T][k] rows of the current run so far (rows 1..i-1 valid) */
  const double* theta0;
  const double* theta1_prev; /* env$theta1: the previous proposal (RAM copies fixed coords from it) */
  double f0;
  int64_t i;    /* 1-based row being proposed */
  int64_t nsteps;
  const fmcmc_model_desc* model;
} loop_env;

/* returns 0 or an FMCMC_E* code */
static int propose(kchain* kc, const draw_src* ds, const loop_env* env, double* prop, char* err,
                   size_t errlen) {
  const fmcmc_kernel_spec* ks = kc->ks;
  const int k = kc->k, kf = kc->kf;
  int act[k > 0 ? k : 1];
  uint8_t whichmask[k > 0 ? k : 1];
  const int64_t i = env->i;
  int64_t* abs_iter = &kc->istate[0];
  int64_t* flags = &kc->istate[1];
  memcpy(prop, env->theta0, sizeof(double) * k);

  switch (ks->type) {
    /* R/kernel_normal.R:65-72 and 149-164 */
    case FMCMC_KERNEL_NORMAL:
    case FMCMC_KERNEL_NORMAL_REFLECTIVE: {
      int na = active_coords(kc, ds, i, act);
      memset(whichmask, 0, k);
      for (int j = 0; j < na; j++) {
        int w = act[j];
        whichmask[w] = 1;
        double z = draw_z(ds, j);
        double inc = (ks->scale[w] == 0.) ? ks->mu[w] : ks->mu[w] + ks->scale[w] * z;
        prop[w] = prop[w] + inc;
      }
      if (ks->type == FMCMC_KERNEL_NORMAL_REFLECTIVE)
        fmcmc_oracle_reflect(k, prop, ks->lb, ks->ub, whichmask);
      return 0;
    }
    /* R/kernel_unif.R:53-57 and 124-135 */
    case FMCMC_KERNEL_UNIF:
    case FMCMC_KERNEL_UNIF_REFLECTIVE: {
      int na = active_coords(kc, ds, i, act);
      memset(whichmask, 0, k);
      for (int j = 0; j < na; j++) {
        int w = act[j];
        whichmask[w] = 1;
        double u = draw_u01(ds, j);
        double a = ks->min_[w], b = ks->max_[w];
        double inc = (a == b) ? a : a + (b - a) * u;
        prop[w] = prop[w] + inc;
      }
      if (ks->type == FMCMC_KERNEL_UNIF_REFLECTIVE)
        fmcmc_oracle_reflect(k, prop, ks->lb, ks->ub, whichmask);
      return 0;
    }
    /* R/kernel_adapt.R:84-182 */
    case FMCMC_KERNEL_ADAPT: {
      double* Sigma = kc->dstate;
      double* Mean_prev = kc->dstate + (size_t)kf * kf;
      if (!(*flags & FMCMC_STATE_INIT)) { /* :87-115, Sigma <- Ik = diag(k)*eps */
        memset(Sigma, 0, sizeof(double) * kf * kf);
        for (int a = 0; a < kf; a++) Sigma[a + a * kf] = ks->eps;
        *flags |= FMCMC_STATE_INIT;
      }
      if (ks->until > (double)*abs_iter && *abs_iter > ks->warmup && i > 2 &&
          (i % ks->freq) == 0) { /* :118 */
        if (ks->bw > 0) { /* :121-125  Sigma <- Sd * (cov(ans[ran, which.]) + Ik) */
          int64_t r0 = i - ks->bw + 1, r1 = i - 1; /* 1-based inclusive */
          if (r0 < 1) {
            set_err(err, errlen, "kernel_adapt: bw window starts before row 1 (reference errors)");
            return FMCMC_EUNSUP;
          }
          int64_t nr = r1 - r0 + 1;
          double Sd = ks->Sd > 0 ? ks->Sd : 5.76 / kf;
          for (int b = 0; b < kf; b++)
            for (int a = 0; a < kf; a++) {
              long double ma = 0, mb = 0;
              for (int64_t r = r0; r <= r1; r++) {
                ma += env->ans[(r - 1) * k + kc->free_idx[a]];
                mb += env->ans[(r - 1) * k + kc->free_idx[b]];
              }
              ma /= nr; mb /= nr;
              long double s = 0;
              for (int64_t r = r0; r <= r1; r++)
                s += (env->ans[(r - 1) * k + kc->free_idx[a]] - ma) *
                     (env->ans[(r - 1) * k + kc->free_idx[b]] - mb);
              Sigma[a + b * kf] = Sd * ((double)(s / (nr - 1)) + (a == b ? ks->eps : 0.0));
            }
        } else {
          if (!(*flags & FMCMC_STATE_HAS_MEAN)) { /* :130-131 colMeans(ans[1:(i-1), which.]) */
            for (int a = 0; a < kf; a++) {
              long double s = 0;
              for (int64_t r = 1; r <= i - 1; r++) s += env->ans[(r - 1) * k + kc->free_idx[a]];
              Mean_prev[a] = (double)(s / (long double)(i - 1));
            }
            *flags |= FMCMC_STATE_HAS_MEAN;
          }
          if (i - ks->freq < 1) {
            set_err(err, errlen, "kernel_adapt: update range starts before row 1 (reference errors)");
            return FMCMC_EUNSUP;
          }
          double t = (double)(*abs_iter - ks->freq); /* :144,152 */
          if (t == 0.0) {
            set_err(err, errlen, "kernel_adapt: t. = abs_iter - freq is zero (reference yields NaN Sigma)");
            return FMCMC_EUNSUP;
          }
          double x[kf], m[kf], Ik[kf * kf];
          memset(Ik, 0, sizeof(Ik));
          for (int a = 0; a < kf; a++) Ik[a + a * kf] = ks->eps;
          for (int64_t jj = 0; jj < ks->freq; jj++) { /* rows (i-freq):(i-1) */
            int64_t r = i - ks->freq + jj; /* 1-based */
            for (int a = 0; a < kf; a++) x[a] = env->ans[(r - 1) * k + kc->free_idx[a]];
            double tj = t + (double)jj;
            mean_rec1(kf, x, Mean_prev, tj, m);
            cov_rec1(kf, x, Sigma, m, Mean_prev, tj, 1e-5, 1.0, Ik); /* :148-156, Sd default 1 */
            memcpy(Mean_prev, m, sizeof(double) * kf);
          }
        }
      }
      *abs_iter += 1; /* :170 */
      /* :173-178  theta0[which.] + MASS::mvrnorm(mu, Sigma) */
      double z[kf], inc[kf];
      for (int a = 0; a < kf; a++) z[a] = draw_z(ds, a);
      if (ks->mvn_method == FMCMC_MVN_EIGEN) {
        double A[kf * kf], V[kf * kf], ev[kf];
        memcpy(A, Sigma, sizeof(A));
        jacobi_eigen(kf, A, ev, V);
        for (int a = 0; a < kf; a++)
          if (ev[a] < -1e-6 * fabs(ev[0])) {
            set_err(err, errlen, "'Sigma' is not positive definite");
            return FMCMC_ENOTPD;
          }
        for (int a = 0; a < kf; a++) {
          double s = 0;
          for (int b = 0; b < kf; b++) s += V[a + b * kf] * sqrt(ev[b] > 0 ? ev[b] : 0.0) * z[b];
          inc[a] = s;
        }
      } else {
        double L[kf * kf];
        if (chol_lower(kf, Sigma, L)) {
          set_err(err, errlen, "'Sigma' is not positive definite");
          return FMCMC_ENOTPD;
        }
        for (int a = 0; a < kf; a++) {
          double s = 0;
          for (int b = 0; b <= a; b++) s += L[a + b * kf] * z[b];
          inc[a] = s;
        }
      }
      memset(whichmask, 0, k);
      for (int a = 0; a < kf; a++) {
        int w = kc->free_idx[a];
        whichmask[w] = 1;
        prop[w] = env->theta0[w] + (ks->mu[w] + inc[a]);
      }
      fmcmc_oracle_reflect(k, prop, ks->lb, ks->ub, whichmask); /* :180 */
      return 0;
    }
    /* R/kernel_ram.R:90-160 */
    case FMCMC_KERNEL_RAM: {
      double* S = kc->dstate;
      if (!(*flags & FMCMC_STATE_INIT)) { /* :114-116 Sigma <- Ik * eps */
        memset(S, 0, sizeof(double) * kf * kf);
        for (int a = 0; a < kf; a++) S[a + a * kf] = ks->eps;
        *flags |= FMCMC_STATE_INIT;
      }
      double U[kf];
      for (int a = 0; a < kf; a++) U[a] = draw_t(ds, a, (double)kf); /* :124 qfun(k) */
      memcpy(prop, env->theta1_prev, sizeof(double) * k);            /* :125 theta1 <- env$theta1 */
      memset(whichmask, 0, k);
      for (int a = 0; a < kf; a++) {
        double s = 0;
        for (int b = 0; b < kf; b++) s += S[a + b * kf] * U[b];
        int w = kc->free_idx[a];
        whichmask[w] = 1;
        prop[w] = env->theta0[w] + s; /* :126 */
      }
      if (ks->until > (double)*abs_iter && *abs_iter > ks->warmup && (i % ks->freq) == 0) { /* :129 */
        double a_n = exp(fmcmc_oracle_logpost(env->model, prop) - env->f0); /* :132 un-reflected */
        if (a_n > 1) a_n = 1;
        if (!isfinite(a_n)) a_n = 0.0;
        double eta = pow((double)i, -2.0 / 3.0) * kf; /* :66 */
        if (eta > 1.0) eta = 1.0;
        double nrm2 = 0;
        for (int a = 0; a < kf; a++) nrm2 += U[a] * U[a];
        double nn = sqrt(nrm2);
        nn = nn * nn; /* norm(rbind(U), "2")^2 */
        double cfac = eta * (a_n - ks->arate);
        /* Sigma %*% (Ik + c UU'/|U|^2) %*% t(Sigma)   :136-139 */
        double Mid[kf * kf], T1[kf * kf], Mx[kf * kf];
        for (int b = 0; b < kf; b++)
          for (int a = 0; a < kf; a++)
            Mid[a + b * kf] = (a == b ? 1.0 : 0.0) + cfac * (U[a] * U[b]) / nn;
        for (int b = 0; b < kf; b++)
          for (int a = 0; a < kf; a++) {
            double s = 0;
            for (int c = 0; c < kf; c++) s += S[a + c * kf] * Mid[c + b * kf];
            T1[a + b * kf] = s;
          }
        for (int b = 0; b < kf; b++)
          for (int a = 0; a < kf; a++) {
            double s = 0;
            for (int c = 0; c < kf; c++) s += T1[a + c * kf] * S[b + c * kf];
            Mx[a + b * kf] = s;
          }
        double L[kf * kf];
        if (chol_lower(kf, Mx, L)) { /* :141-146; nearPD is third-party => jitter repair (UNPINNED) */
          kc->istate[2] += 1;
          double md = 0;
          for (int a = 0; a < kf; a++) md += fabs(Mx[a + a * kf]);
          md = md / kf;
          if (!(md > 0) || !isfinite(md)) md = 1.0;
          for (int b = 0; b < kf; b++)
            for (int a = b + 1; a < kf; a++) {
              double v = 0.5 * (Mx[a + b * kf] + Mx[b + a * kf]);
              Mx[a + b * kf] = Mx[b + a * kf] = v;
            }
          double jit = 1e-8 * md;
          int ok = 0;
          for (int tr = 0; tr < 20 && !ok; tr++, jit *= 10) {
            double Mj[kf * kf];
            memcpy(Mj, Mx, sizeof(Mj));
            for (int a = 0; a < kf; a++) Mj[a + a * kf] += jit;
            ok = !chol_lower(kf, Mj, L);
          }
          if (!ok) {
            set_err(err, errlen, "kernel_ram: Sigma could not be repaired");
            return FMCMC_ENOTPD;
          }
        }
        memcpy(S, L, sizeof(double) * kf * kf);
        if (ks->constr) /* :149-150 */
          for (int b = 0; b < kf; b++)
            for (int a = 0; a < kf; a++)
              S[a + b * kf] *= ks->constr[kc->free_idx[a] + (size_t)kc->free_idx[b] * k];
      }
      *abs_iter += 1;                                            /* :155 */
      fmcmc_oracle_reflect(k, prop, ks->lb, ks->ub, whichmask); /* :158 */
      return 0;
    }
    /* R/kernel_mirror.R:74-157 (nmirror) and 198-284 (umirror) */
    case FMCMC_KERNEL_NMIRROR:
    case FMCMC_KERNEL_UMIRROR: {
      double* mu = kc->dstate;
      double* scale = kc->dstate + k;
      double* obs = kc->dstate + 2 * k;
      if (!(*flags & FMCMC_STATE_INIT)) {
        for (int j = 0; j < k; j++) { mu[j] = ks->mu[j]; scale[j] = ks->scale[j]; obs[j] = 0; }
        *flags |= FMCMC_STATE_INIT;
      }
      const int64_t ai = *abs_iter;
      const int64_t nad0 = ks->nadapt_len > 0 ? ks->nadapt[0] : -1;
      if (ai >= 1 && ai <= ks->warmup) { /* :104-112 mean_recursive on ans[i-1,] */
        const double* x = env->ans + (i - 2) * k;
        for (int j = 0; j < k; j++) mu[j] = (mu[j] * (double)ai + x[j]) / ((double)ai + 1);
      }
      if (ai == nad0) { /* :115-119 */
        if (i - 1 < 2) {
          set_err(err, errlen, "mirror kernel: acceptance rate needs >= 2 rows (quirk D8; reference yields NaN)");
          return FMCMC_EUNSUP;
        }
        int64_t same = 0;
        for (int64_t r = 2; r <= i - 1; r++) {
          double s = 0;
          for (int j = 0; j < k; j++) {
            double dd = env->ans[(r - 1) * k + j] - env->ans[(r - 2) * k + j];
            s += dd * dd;
          }
          if (s == 0.0) same++;
        }
        double v = 1.0 - (double)same / (double)(i - 2);
        for (int j = 0; j < k; j++) obs[j] = v;
        *flags = (*flags & ~(3 << FMCMC_STATE_OBS_SHIFT)) | (1 << FMCMC_STATE_OBS_SHIFT);
      } else if (ai > nad0 && ai <= ks->warmup && nad0 >= 0) { /* :120-128 */
        if (i - 2 < 1) {
          set_err(err, errlen, "mirror kernel: ans[i-2,] does not exist at i = 2 (quirk D8); use freq >= warmup + 2");
          return FMCMC_EUNSUP;
        }
        const double* x1 = env->ans + (i - 2) * k;
        const double* x2 = env->ans + (i - 3) * k;
        for (int j = 0; j < k; j++) {
          double ind = (x1[j] != x2[j]) ? 1.0 : 0.0;
          obs[j] = (obs[j] * (double)ai + ind) / ((double)ai + 1);
        }
        *flags = (*flags & ~(3 << FMCMC_STATE_OBS_SHIFT)) | (2 << FMCMC_STATE_OBS_SHIFT);
      }
      int in_nadapt = 0;
      for (int q = 0; q < ks->nadapt_len; q++)
        if (ks->nadapt[q] == ai) in_nadapt = 1;
      if (in_nadapt) /* :131-137 */
        for (int j = 0; j < k; j++)
          scale[j] = scale[j] * tan(M_PI / 2.0 * obs[j]) / tan(M_PI / 2.0 * ks->arate);
      int na = active_coords(kc, ds, i, act);
      memset(whichmask, 0, k);
      const double sqrt3 = sqrt(3.0);
      for (int j = 0; j < na; j++) {
        int w = act[j];
        whichmask[w] = 1;
        if (ks->type == FMCMC_KERNEL_NMIRROR) { /* :146-150 */
          double mean = 2 * mu[w] - env->theta0[w];
          double z = draw_z(ds, j);
          prop[w] = (scale[w] == 0.) ? mean : mean + scale[w] * z;
        } else { /* :273-277 (joint & no fixed only: D11 rejected upstream) */
          double a = 2 * mu[w] - env->theta0[w] - sqrt3 * scale[w];
          double b = 2 * mu[w] - env->theta0[w] + sqrt3 * scale[w];
          double u = draw_u01(ds, j);
          prop[w] = (a == b) ? a : a + (b - a) * u;
        }
      }
      *abs_iter += 1;
      fmcmc_oracle_reflect(k, prop, ks->lb, ks->ub, whichmask);
      return 0;
    }
  }
  set_err(err, errlen, "unknown kernel type %d", ks->type);
  return FMCMC_EINVAL;
}

/* ======================================================================== */
/* R/mcmc.R:485-838  MCMC_without_conv_checker (serial chain fan-out 643-673, */
/* loop 726-783, burnin/thin 786-813)                                         */
/* ======================================================================== */
int64_t fmcmc_oracle_rows_kept(int64_t nsteps, int64_t burnin, int64_t thin) {
  int64_t m = nsteps - burnin;
  if (m < 0) return 0;
  if (thin < 1) thin = 1;
  return m / thin; /* which((1:m %% thin) == 0), R/mcmc.R:803 */
}

static int validate(const fmcmc_run_spec* run, const fmcmc_kernel_spec* ks, int k, int* kf_out,
                    char* err, size_t errlen) {
  if (run->nchains < 1) { set_err(err, errlen, "`nchains` must be an integer greater than 1."); return FMCMC_EINVAL; }
  if (run->burnin >= run->nsteps) {
    set_err(err, errlen, "-burnin- (%lld) cannot be >= than -nsteps- (%lld).", (long long)run->burnin, (long long)run->nsteps);
    return FMCMC_EINVAL;
  }
  if (run->thin >= run->nsteps) {
    set_err(err, errlen, "-thin- (%lld) cannot be > than -nsteps- (%lld).", (long long)run->thin, (long long)run->nsteps);
    return FMCMC_EINVAL;
  }
  if (run->thin < 1) { set_err(err, errlen, "-thin- should be >= 1."); return FMCMC_EINVAL; }
  if (ks->k != k) { set_err(err, errlen, "Incorrect length of -initial-: kernel k=%d, family k=%d.", ks->k, k); return FMCMC_EINVAL; }
  int kf = 0;
  for (int j = 0; j < k; j++) kf += ks->fixed && ks->fixed[j] ? 0 : 1;
  if (kf == 0) {
    set_err(err, errlen, "The number of parameters to update, i.e. not fixed, cannot be zero. Check the value -fixed- in the kernel initialization.");
    return FMCMC_EINVAL;
  }
  int bounded = !(ks->type == FMCMC_KERNEL_NORMAL || ks->type == FMCMC_KERNEL_UNIF);
  if (bounded)
    for (int j = 0; j < k; j++)
      if (ks->ub[j] <= ks->lb[j]) { set_err(err, errlen, "-ub- cannot be <= than -lb-."); return FMCMC_EINVAL; }
  if (ks->type == FMCMC_KERNEL_UNIF || ks->type == FMCMC_KERNEL_UNIF_REFLECTIVE)
    for (int j = 0; j < k; j++)
      if (ks->max_[j] <= ks->min_[j]) { set_err(err, errlen, "-max.- cannot be <= than -min.-."); return FMCMC_EINVAL; }
  if (ks->type == FMCMC_KERNEL_UMIRROR && (kf != k || ks->scheme != FMCMC_SCHEME_JOINT)) {
    set_err(err, errlen, "kernel_umirror with fixed parameters or a non-joint scheme is ill-defined in the reference (quirk D11)");
    return FMCMC_EUNSUP;
  }
  if (ks->type == FMCMC_KERNEL_ADAPT && ks->bw > 0 && ks->bw > ks->warmup) {
    set_err(err, errlen, "The `warmup` parameter must be greater than `bw`.");
    return FMCMC_EINVAL;
  }
  if (ks->scheme == FMCMC_SCHEME_EXPLICIT) {
    if (ks->order_len != kf) {
      set_err(err, errlen, "When setting the update scheme, it should have the same length as the number of variables that will not be fixed.");
      return FMCMC_EINVAL;
    }
  }
  *kf_out = kf;
  return 0;
}

static int fmcmc_oracle_threads = 1;
void fmcmc_oracle_set_threads(int n) { fmcmc_oracle_threads = n < 1 ? 1 : n; }

typedef struct {
  const fmcmc_model_desc* model; const fmcmc_run_spec* run; const fmcmc_kernel_spec* ks;
  fmcmc_kernel_state* state; const fmcmc_stream_spec* stream;
  double *ans_out, *draws_out, *logpost_out; fmcmc_run_report* report; char* err; size_t errlen;
  int k, kf; int64_t T, keep, dlen; int* free_idx; int kdraw;
  int status; int64_t n_accept; int next_chain; pthread_mutex_t mu;
} run_ctx;

static void run_chain(run_ctx* cx, int c) {
  const fmcmc_model_desc* model = cx->model; const fmcmc_run_spec* run = cx->run;
  const fmcmc_kernel_spec* ks = cx->ks; fmcmc_kernel_state* state = cx->state;
  const fmcmc_stream_spec* stream = cx->stream;
  double *ans_out = cx->ans_out, *draws_out = cx->draws_out, *logpost_out = cx->logpost_out;
  const int k = cx->k, kf = cx->kf, kdraw = cx->kdraw; const int64_t T = cx->T, keep = cx->keep, dlen = cx->dlen;
  int* free_idx = cx->free_idx;
  int64_t my_accept = 0;
    if (__atomic_load_n(&cx->status, __ATOMIC_RELAXED)) return;
    double* ans = (double*)malloc(sizeof(double) * T * k);
    double* draws = (double*)malloc(sizeof(double) * T * k);
    double* logpost = (double*)malloc(sizeof(double) * T);
    double theta0[k], theta1[k], prop[k];
    int64_t istate_local[FMCMC_ISTATE_LEN] = {0, 0, 0, 0};
    kchain kc;
    kc.ks = ks; kc.k = k; kc.kf = kf; kc.free_idx = free_idx;
    kc.istate = state && state->istate ? state->istate + (size_t)c * FMCMC_ISTATE_LEN : istate_local;
    kc.dstate = state && state->dstate ? state->dstate + (size_t)c * dlen : NULL;
    double* dtmp = NULL;
    if (!kc.dstate && dlen) { dtmp = (double*)calloc(dlen, sizeof(double)); kc.dstate = dtmp; }
    kc.seq = (ks->scheme == FMCMC_SCHEME_RANDOM && ks->seq) ? ks->seq + (size_t)c * ks->seq_len : NULL;
    kc.istate[3] = 0;
    char lerr[256] = {0};
    int lrc = 0;

    /* R/mcmc.R:737-743 */
    memcpy(theta0, run->initial + (size_t)c * k, sizeof(double) * k);
    memcpy(theta1, theta0, sizeof(double) * k);
    memcpy(ans, theta0, sizeof(double) * k);
    memcpy(draws, theta0, sizeof(double) * k);
    double f0 = fmcmc_oracle_logpost(model, theta0);
    logpost[0] = f0;

    draw_src ds;
    ds.mode = stream->mode; ds.seed = stream->seed;
    ds.chain = (uint32_t)(run->chain_offset + c); ds.run = (uint32_t)stream->run_index;
    loop_env env;
    env.ans = ans; env.nsteps = T; env.model = model;

    int64_t i;
    for (i = 2; i <= T; i++) { /* R/mcmc.R:749-783 */
      ds.row = (uint32_t)i;
      ds.zrow = (stream->mode == FMCMC_STREAM_FED)
                    ? stream->z + ((size_t)c * T + (size_t)(i - 1)) * kdraw : NULL;
      env.theta0 = theta0; env.theta1_prev = theta1; env.f0 = f0; env.i = i;
      lrc = propose(&kc, &ds, &env, prop, lerr, sizeof(lerr));
      if (lrc) break;
      memcpy(draws + (i - 1) * k, prop, sizeof(double) * k);
      memcpy(theta1, prop, sizeof(double) * k);
      double f1 = fmcmc_oracle_logpost(model, theta1);
      logpost[i - 1] = f1;
      if (isnan(f1)) { /* :758-765 */
        snprintf(lerr, sizeof(lerr),
                 "fun(par) is undefined (NaN). Check either -fun- or the -lb- and -ub- parameters. "
                 "This error ocurred during step i = %lld (chain %d).", (long long)i, c + 1);
        lrc = FMCMC_ENAN;
        break;
      }
      double klogratio = f1 - f0; /* R/kernel.R:302-303 */
      if (isnan(klogratio)) {     /* D10 */
        snprintf(lerr, sizeof(lerr), "missing value where TRUE/FALSE needed (f1 - f0 is NaN) at step i = %lld (chain %d).", (long long)i, c + 1);
        lrc = FMCMC_ENANRATIO;
        break;
      }
      double logu;
      if (stream->mode == FMCMC_STREAM_FED) {
        logu = stream->logu[(size_t)c * T + (size_t)(i - 1)];
      } else {
        double u0, u1;
        fmcmc_oracle_philox_u2(ds.seed, ds.chain, ds.run, ds.row, 0u, &u0, &u1);
        logu = log(u0);
      }
      if (logu < klogratio) { /* :770 */
        int changed = 0;
        for (int j = 0; j < k; j++) changed |= (theta0[j] != theta1[j]);
        kc.istate[3] += changed;
        memcpy(theta0, theta1, sizeof(double) * k);
        f0 = f1;
        my_accept += 1;
      }
      memcpy(ans + (i - 1) * k, theta0, sizeof(double) * k); /* :778 */
    }
    if (lrc) {
      pthread_mutex_lock(&cx->mu);
      if (!cx->status) {
        cx->status = lrc;
        set_err(cx->err, cx->errlen, "%s", lerr);
        if (cx->report) { cx->report->nan_chain = c + 1; cx->report->nan_step = i; }
      }
      pthread_mutex_unlock(&cx->mu);
    } else {
      /* burnin + thin, R/mcmc.R:786-813: keep post-burnin positions p with p %% thin == 0 */
      for (int64_t r = 0; r < keep; r++) {
        int64_t src = run->burnin + (r + 1) * run->thin - 1; /* 0-based row */
        if (ans_out) memcpy(ans_out + ((size_t)c * keep + r) * k, ans + src * k, sizeof(double) * k);
        if (draws_out) memcpy(draws_out + ((size_t)c * keep + r) * k, draws + src * k, sizeof(double) * k);
        if (logpost_out) logpost_out[(size_t)c * keep + r] = logpost[src];
      }
    }
    free(ans); free(draws); free(logpost); free(dtmp);
    __atomic_fetch_add(&cx->n_accept, my_accept, __ATOMIC_RELAXED);
  }

static void* run_worker(void* p) {
  run_ctx* cx = (run_ctx*)p;
  for (;;) {
    int c = __atomic_fetch_add(&cx->next_chain, 1, __ATOMIC_RELAXED);
    if (c >= cx->run->nchains) break;
    run_chain(cx, c);
  }
  return NULL;
}

int fmcmc_oracle_run(const fmcmc_model_desc* model, const fmcmc_run_spec* run,
                     const fmcmc_kernel_spec* ks, fmcmc_kernel_state* state,
                     const fmcmc_stream_spec* stream, double* ans_out, double* draws_out,
                     double* logpost_out, fmcmc_run_report* report, char* err, size_t errlen) {
  const int k = fmcmc_oracle_nparams(model);
  int kf = 0;
  int rc = validate(run, ks, k, &kf, err, errlen);
  if (rc) return rc;
  const int64_t T = run->nsteps;
  const int64_t keep = fmcmc_oracle_rows_kept(T, run->burnin, run->thin);
  const int64_t dlen = fmcmc_oracle_state_len(ks->type, k, kf);
  int free_idx[k];
  {
    int a = 0;
    for (int j = 0; j < k; j++)
      if (!(ks->fixed && ks->fixed[j])) free_idx[a++] = j;
  }
  if (report) {
    memset(report, 0, sizeof(*report));
    report->rows_kept = keep;
    report->first_iter = run->burnin + run->thin;
    report->last_iter = run->burnin + keep * run->thin;
  }
  int status = 0;
  int64_t n_accept = 0;
  const int kdraw = stream->kdraw;

  run_ctx cx;
  memset(&cx, 0, sizeof(cx));
  cx.model = model; cx.run = run; cx.ks = ks; cx.state = state; cx.stream = stream;
  cx.ans_out = ans_out; cx.draws_out = draws_out; cx.logpost_out = logpost_out;
  cx.report = report; cx.err = err; cx.errlen = errlen;
  cx.k = k; cx.kf = kf; cx.T = T; cx.keep = keep; cx.dlen = dlen; cx.free_idx = free_idx; cx.kdraw = kdraw;
  pthread_mutex_init(&cx.mu, NULL);
  int nthreads = fmcmc_oracle_threads;
  if (nthreads > run->nchains) nthreads = run->nchains;
  if (nthreads <= 1) {
    run_worker(&cx);
  } else { /* one chain per worker at a time: the PSOCK decomposition, R/mcmc.R:593-627 */
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, run_worker, &cx);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
  }
  pthread_mutex_destroy(&cx.mu);
  status = cx.status; n_accept = cx.n_accept;
  if (report) report->n_accept = n_accept;
  return status;
}

/* ======================================================================== */
/* coda::gelman.diag as called at R/convergence.R:207 (third-party: coda, no  */
/* version pin in DESCRIPTION:36-42; defaults confidence=.95, transform=FALSE,*/
/* autoburnin=TRUE (windowing done by the caller), multivariate=TRUE).        */
/* x: [m chains][N rows][p vars] row-major, already windowed.                 */
/* ======================================================================== */
int fmcmc_oracle_gelman(int64_t m, int64_t N, int p, const double* x, double* psrf,
                        double* mpsrf) {
  double* xbar = (double*)calloc((size_t)m * p, sizeof(double));
  double* S2 = (double*)calloc((size_t)m * p * p, sizeof(double));
  double* W = (double*)calloc((size_t)p * p, sizeof(double));
  double* B = (double*)calloc((size_t)p * p, sizeof(double));
  int rc = 0;
  for (int64_t j = 0; j < m; j++) {
    const double* xj = x + (size_t)j * N * p;
    for (int a = 0; a < p; a++) {
      long double s = 0;
      for (int64_t t = 0; t < N; t++) s += xj[t * p + a];
      long double mean = s / N;
      long double tt = 0;
      for (int64_t t = 0; t < N; t++) tt += xj[t * p + a] - mean;
      mean += tt / N;
      xbar[j * p + a] = (double)mean;
    }
    for (int a = 0; a < p; a++)
      for (int b = 0; b <= a; b++) {
        long double s = 0;
        for (int64_t t = 0; t < N; t++)
          s += ((long double)xj[t * p + a] - xbar[j * p + a]) * ((long double)xj[t * p + b] - xbar[j * p + b]);
        double v = (double)(s / (N - 1));
        S2[(size_t)j * p * p + a + b * p] = v;
        S2[(size_t)j * p * p + b + a * p] = v;
      }
  }
  for (int e = 0; e < p * p; e++) {
    long double s = 0;
    for (int64_t j = 0; j < m; j++) s += S2[(size_t)j * p * p + e];
    W[e] = (double)(s / m);
  }
  double muhat[p];
  for (int a = 0; a < p; a++) {
    long double s = 0;
    for (int64_t j = 0; j < m; j++) s += xbar[j * p + a];
    muhat[a] = (double)(s / m);
  }
  for (int a = 0; a < p; a++)
    for (int b = 0; b < p; b++) {
      long double s = 0;
      for (int64_t j = 0; j < m; j++) s += (xbar[j * p + a] - muhat[a]) * (xbar[j * p + b] - muhat[b]);
      B[a + b * p] = (double)N * (double)(s / (m - 1));
    }
  *mpsrf = NAN;
  if (p > 1) {
    double* CW = (double*)calloc((size_t)p * p, sizeof(double)); /* lower L, W = L L' (CW = L') */
    if (chol_lower(p, W, CW)) {
      rc = FMCMC_ENOTPD;
    } else {
      /* M = L^{-1} B L^{-T} (= backsolve(CW, t(backsolve(CW, B, transpose=TRUE)), transpose=TRUE)) */
      double* Y = (double*)calloc((size_t)p * p, sizeof(double));
      double* Mx = (double*)calloc((size_t)p * p, sizeof(double));
      for (int c = 0; c < p; c++) /* solve L Y = B */
        for (int a = 0; a < p; a++) {
          double s = B[a + c * p];
          for (int q = 0; q < a; q++) s -= CW[a + q * p] * Y[q + c * p];
          Y[a + c * p] = s / CW[a + a * p];
        }
      for (int c = 0; c < p; c++) /* solve L Mx' = Y'  => Mx = Y L^{-T} */
        for (int a = 0; a < p; a++) {
          double s = Y[c + a * p];
          for (int q = 0; q < a; q++) s -= CW[a + q * p] * Mx[q + c * p];
          Mx[a + c * p] = s / CW[a + a * p];
        }
      for (int a = 0; a < p; a++)
        for (int b = 0; b < a; b++) {
          double v = 0.5 * (Mx[a + b * p] + Mx[b + a * p]);
          Mx[a + b * p] = Mx[b + a * p] = v;
        }
      double* ev = (double*)calloc(p, sizeof(double));
      double* V = (double*)calloc((size_t)p * p, sizeof(double));
      jacobi_eigen(p, Mx, ev, V);
      double emax = ev[0];
      *mpsrf = sqrt((1.0 - 1.0 / (double)N) + (1.0 + 1.0 / (double)p) * emax / (double)N);
      free(Y); free(Mx); free(ev); free(V);
    }
    free(CW);
  }
  /* univariate psrf point estimates */
  for (int a = 0; a < p; a++) {
    double w = W[a + a * p], b = B[a + a * p];
    long double ms2 = 0, mx = 0, mx2 = 0;
    for (int64_t j = 0; j < m; j++) {
      ms2 += S2[(size_t)j * p * p + a + a * p];
      mx += xbar[j * p + a];
      mx2 += xbar[j * p + a] * xbar[j * p + a];
    }
    ms2 /= m; mx /= m; mx2 /= m;
    long double vs2 = 0, c1 = 0, c2 = 0;
    for (int64_t j = 0; j < m; j++) {
      long double ds2 = S2[(size_t)j * p * p + a + a * p] - ms2;
      vs2 += ds2 * ds2;
      c1 += ds2 * (xbar[j * p + a] * xbar[j * p + a] - mx2);
      c2 += ds2 * (xbar[j * p + a] - mx);
    }
    double var_w = (double)(vs2 / (m - 1)) / (double)m;
    double var_b = (2.0 * b * b) / (double)(m - 1);
    double cov_wb = ((double)N / (double)m) * ((double)(c1 / (m - 1)) - 2.0 * muhat[a] * (double)(c2 / (m - 1)));
    double Vv = ((double)N - 1) * w / (double)N + (1.0 + 1.0 / (double)m) * b / (double)N;
    double var_V = (((double)N - 1) * ((double)N - 1) * var_w + (1.0 + 1.0 / (double)m) * (1.0 + 1.0 / (double)m) * var_b +
                    2.0 * ((double)N - 1) * (1.0 + 1.0 / (double)m) * cov_wb) / ((double)N * (double)N);
    double df_V = (2.0 * Vv * Vv) / var_V;
    double df_adj = (df_V + 3) / (df_V + 1);
    double R2_fixed = ((double)N - 1) / (double)N;
    double R2_random = (1.0 + 1.0 / (double)m) * (1.0 / (double)N) * (b / w);
    psrf[a] = sqrt(df_adj * (R2_fixed + R2_random));
  }
  free(xbar); free(S2); free(W); free(B);
  return rc;
}

/* R/convergence.R:169-186 rm_invariant: ONE pooled scalar variance (quirk D9).
 * Returns 1 when the pooled variance of all values is < 1e-10. */
int fmcmc_oracle_pooled_invariant(int64_t count, const double* x) {
  long double s = 0;
  for (int64_t i = 0; i < count; i++) s += x[i];
  long double mean = s / count, ss = 0;
  for (int64_t i = 0; i < count; i++) ss += (x[i] - mean) * (x[i] - mean);
  double sd = sqrt((double)(ss / (count - 1)));
  return sd * sd < 1e-10;
}

int fmcmc_oracle_version(void) { return FMCMC_ABI_VERSION; }
