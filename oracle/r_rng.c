/*
 * oracle/r_rng.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A restatement of the parts of base R's random-number machinery that fmcmc's
 * serial path consumes, so that the golden outputs
 * printed in the reference's README (README.md:183-201, 315-339, 388-412) can
 * be regenerated without an R installation and used to PIN the oracle.
 *
 * Third-party algorithm, not present under /root/reference: base R >= 3.3
 * (DESCRIPTION:22), default RNGkind("Mersenne-Twister", "Inversion"):
 *   - set.seed(): 50 rounds of the LCG 69069*s+1, then 625 more to fill the
 *     state, then mti = 624 (R: src/main/RNG.c RNG_Init / FixupSeeds);
 *   - unif_rand(): MT19937 genrand * 2.3283064365386963e-10, clamped into (0,1);
 *   - norm_rand() INVERSION: u = unif; u = (int)(2^27 u) + unif; qnorm(u / 2^27);
 *   - qnorm(): Wichura's AS 241 (PPND16) as in R's nmath/qnorm.c;
 *   - rnorm(mu, s) = mu + s*norm_rand(); runif(a, b) = a + (b-a)*unif_rand().
 * Checked in tests/test_oracle_rrng.py against numpy's MT19937 (same state =>
 * same 32-bit outputs) and scipy's norm.ppf.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MT_N 624
#define MT_M 397

static uint32_t g_dummy[MT_N + 1]; /* [0] = mti, [1..624] = mt, like R's .Random.seed[-1] */

void r_set_seed(uint32_t seed) {
  for (int j = 0; j < 50; j++) seed = 69069u * seed + 1u;
  for (int j = 0; j < MT_N + 1; j++) {
    seed = 69069u * seed + 1u;
    g_dummy[j] = seed;
  }
  g_dummy[0] = MT_N; /* FixupSeeds: mti = N */
}

/* expose / restore the state (for the numpy cross-check) */
void r_get_state(uint32_t* out625) { memcpy(out625, g_dummy, sizeof(g_dummy)); }
void r_put_state(const uint32_t* in625) { memcpy(g_dummy, in625, sizeof(g_dummy)); }

static uint32_t mt_next_u32(void) {
  static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
  uint32_t* mt = g_dummy + 1;
  uint32_t y;
  int mti = (int)g_dummy[0];
  if (mti >= MT_N) {
    int kk;
    for (kk = 0; kk < MT_N - MT_M; kk++) {
      y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ mag01[y & 0x1u];
    }
    for (; kk < MT_N - 1; kk++) {
      y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 0x1u];
    }
    y = (mt[MT_N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ mag01[y & 0x1u];
    mti = 0;
  }
  y = mt[mti++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  g_dummy[0] = (uint32_t)mti;
  return y;
}

uint32_t r_mt_u32(void) { return mt_next_u32(); }

double r_unif_rand(void) {
  const double i2_32m1 = 2.328306437080797e-10; /* 1/(2^32 - 1) */
  double value = (double)mt_next_u32() * 2.3283064365386963e-10;
  if (value <= 0.0) return 0.5 * i2_32m1;
  if ((1.0 - value) <= 0.0) return 1.0 - 0.5 * i2_32m1;
  return value;
}

/* AS 241 / R nmath/qnorm.c, lower tail, non-log, mu = 0, sigma = 1 */
double r_qnorm(double p) {
  double q, r, val;
  if (isnan(p)) return p;
  if (p <= 0.0) return (p == 0.0) ? -INFINITY : NAN;
  if (p >= 1.0) return (p == 1.0) ? INFINITY : NAN;
  q = p - 0.5;
  if (fabs(q) <= 0.425) {
    r = .180625 - q * q;
    val = q * (((((((r * 2509.0809287301226727 +
                     33430.575583588128105) * r + 67265.770927008700853) * r +
                   45921.953931549871457) * r + 13731.693765509461125) * r +
                 1971.5909503065514427) * r + 133.14166789178437745) * r +
               3.387132872796366608) /
          (((((((r * 5226.495278852545925 +
                 28729.085735721942674) * r + 39307.89580009271061) * r +
               21213.794301586595867) * r + 5394.1960214247511077) * r +
             687.1870074920579083) * r + 42.313330701600911252) * r + 1.);
    return val;
  }
  r = (q < 0) ? p : 1.0 - p;
  r = sqrt(-log(r));
  if (r <= 5.) {
    r += -1.6;
    val = (((((((r * 7.7454501427834140764e-4 +
                 .0227238449892691845833) * r + .24178072517745061177) * r +
               1.27045825245236838258) * r + 3.64784832476320460504) * r +
             5.7694972214606914055) * r + 4.6303378461565452959) * r +
           1.42343711074968357734) /
          (((((((r * 1.05075007164441684324e-9 +
                 5.475938084995344946e-4) * r + .0151986665636164571966) * r +
               .14810397642748007459) * r + .68976733498510000455) * r +
             1.6763848301838038494) * r + 2.05319162663775882187) * r + 1.);
  } else {
    r += -5.;
    val = (((((((r * 2.01033439929228813265e-7 +
                 2.71155556874348757815e-5) * r + .0012426609473880784386) * r +
               .026532189526576123093) * r + .29656057182850489123) * r +
             1.7848265399172913358) * r + 5.4637849111641143699) * r +
           6.6579046435011037772) /
          (((((((r * 2.04426310338993978564e-15 +
                 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r +
               7.868691311456132591e-4) * r + .0148753612908506148525) * r +
             .13692988092273580531) * r + .59983220655588793769) * r + 1.);
  }
  if (q < 0.0) val = -val;
  return val;
}

double r_norm_rand(void) {
  const double BIG = 134217728.0; /* 2^27 */
  double u = r_unif_rand();
  u = (int)(BIG * u) + r_unif_rand();
  return r_qnorm(u / BIG);
}

double r_rnorm1(double mu, double sigma) {
  if (isnan(mu) || !isfinite(sigma) || sigma < 0.) return NAN;
  if (sigma == 0. || !isfinite(mu)) return mu;
  return mu + sigma * r_norm_rand();
}

double r_runif1(double a, double b) {
  if (!isfinite(a) || !isfinite(b) || b < a) return NAN;
  if (a == b) return a;
  double u;
  do { u = r_unif_rand(); } while (u <= 0 || u >= 1);
  return a + (b - a) * u;
}

void r_rnorm(int64_t n, double mu, double sigma, double* out) {
  for (int64_t i = 0; i < n; i++) out[i] = r_rnorm1(mu, sigma);
}
void r_runif(int64_t n, double a, double b, double* out) {
  for (int64_t i = 0; i < n; i++) out[i] = r_runif1(a, b);
}
/* standard normals, i.e. the z that rnorm(mu, s) multiplies (fed-stream upload) */
void r_norm_rand_vec(int64_t n, double* out) {
  for (int64_t i = 0; i < n; i++) out[i] = r_norm_rand();
}
/* log(runif(n)) as R/mcmc.R:726 */
void r_log_runif(int64_t n, double* out) {
  for (int64_t i = 0; i < n; i++) out[i] = log(r_runif1(0.0, 1.0));
}

/* stats::sd() — cov.c: long-double mean, then long-double sum of squares / (n-1) */
double r_sd(const double* x, int64_t n) {
  long double s = 0.0L;
  for (int64_t i = 0; i < n; i++) s += x[i];
  long double m = s / n;
  /* R's cov.c refines the mean once (MEAN macro: second pass of residuals) */
  long double t = 0.0L;
  for (int64_t i = 0; i < n; i++) t += (x[i] - m);
  m += t / n;
  long double ss = 0.0L;
  for (int64_t i = 0; i < n; i++) {
    long double d = x[i] - m;
    ss += d * d;
  }
  return sqrt((double)(ss / (n - 1)));
}

/* ------------------------------------------------------------------------ */
/* exp_rand / rgamma / rchisq / rt — what kernel_ram's default qfun consumes */
/* (R/kernel_ram.R:68 `stats::rt(k, k)`, called at :124).  Third-party, base  */
/* R's nmath (sexp.c, rgamma.c, rchisq.c, rt.c), restated from the published  */
/* algorithms: Ahrens & Dieter (1972) SA for the exponential, Ahrens & Dieter */
/* (1982) GD for shape >= 1, Ahrens & Dieter (1974) GS for shape < 1.  Pinned */
/* by README.md:245-246 (tests/test_oracle_readme_golden.py).                 */
/* ------------------------------------------------------------------------ */
double r_exp_rand(void) {
  /* q[k-1] = sum_{j=1..k} log(2)^j / j! */
  static const double q[] = {
      0.6931471805599453, 0.9333736875190459, 0.9888777961838675, 0.9984589039328340,
      0.9998292811061389, 0.9999833164100727, 0.9999985691438767, 0.9999998906925558,
      0.9999999924734159, 0.9999999995283275, 0.9999999999728814, 0.9999999999985598,
      0.9999999999999289, 0.9999999999999968, 0.9999999999999999, 1.0000000000000000};
  double a = 0.;
  double u = r_unif_rand();
  while (u <= 0. || u >= 1.) u = r_unif_rand();
  for (;;) {
    u += u;
    if (u > 1.) break;
    a += q[0];
  }
  u -= 1.;
  if (u <= q[0]) return a + u;
  int i = 0;
  double ustar = r_unif_rand(), umin = ustar;
  do {
    ustar = r_unif_rand();
    if (umin > ustar) umin = ustar;
    i++;
  } while (u > q[i]);
  return a + umin * q[0];
}
void r_exp_rand_q(double* out16) { /* the table, recomputed, for the unit test */
  long double s = 0, term = 1, l2 = 0.693147180559945309417232121458L;
  for (int k = 1; k <= 16; k++) {
    term *= l2 / k;
    s += term;
    out16[k - 1] = (double)s;
  }
}

double r_rgamma(double a, double scale) {
  const double sqrt32 = 5.656854;
  const double exp_m1 = 0.36787944117144233;
  const double q1 = 0.04166669, q2 = 0.02083148, q3 = 0.00801191, q4 = 0.00144121,
               q5 = -7.388e-5, q6 = 2.4511e-4, q7 = 2.424e-4;
  const double a1 = 0.3333333, a2 = -0.250003, a3 = 0.2000062, a4 = -0.1662921,
               a5 = 0.1423657, a6 = -0.1367177, a7 = 0.1233795;
  static double aa = 0., aaa = 0.;
  static double s, s2, d;
  static double q0, b, si, c;
  double e, p, q, r, t, u, v, w, x, ret_val;

  if (isnan(a) || isnan(scale)) return NAN;
  if (a <= 0.0 || scale <= 0.0) {
    if (scale == 0. || a == 0.) return 0.;
    return NAN;
  }
  if (!isfinite(a) || !isfinite(scale)) return INFINITY;

  if (a < 1) { /* GS */
    e = 1.0 + exp_m1 * a;
    for (;;) {
      p = e * r_unif_rand();
      if (p >= 1.0) {
        x = -log((e - p) / a);
        if (r_exp_rand() >= (1.0 - a) * log(x)) break;
      } else {
        x = exp(log(p) / a);
        if (r_exp_rand() >= x) break;
      }
    }
    return scale * x;
  }

  /* GD, step 1 */
  if (a != aa) {
    aa = a;
    s2 = a - 0.5;
    s = sqrt(s2);
    d = sqrt32 - s * 12;
  }
  /* step 2: immediate acceptance */
  t = r_norm_rand();
  x = s + 0.5 * t;
  ret_val = x * x;
  if (t >= 0) return scale * ret_val;
  /* step 3: squeeze acceptance */
  u = r_unif_rand();
  if (d * u <= t * t * t) return scale * ret_val;
  /* step 4 */
  if (a != aaa) {
    aaa = a;
    r = 1 / a;
    q0 = ((((((q7 * r + q6) * r + q5) * r + q4) * r + q3) * r + q2) * r + q1) * r;
    if (a <= 3.686) {
      b = 0.463 + s + 0.178 * s2;
      si = 1.235;
      c = 0.195 / s - 0.079 + 0.16 * s;
    } else if (a <= 13.022) {
      b = 1.654 + 0.0076 * s2;
      si = 1.68 / s + 0.275;
      c = 0.062 / s + 0.024;
    } else {
      b = 1.77;
      si = 0.75;
      c = 0.1515 / s;
    }
  }
  /* step 5-7 */
  if (x > 0.0) {
    v = t / (s + s);
    if (fabs(v) <= 0.25)
      q = q0 + 0.5 * t * t * ((((((a7 * v + a6) * v + a5) * v + a4) * v + a3) * v + a2) * v + a1) * v;
    else
      q = q0 - s * t + 0.25 * t * t + (s2 + s2) * log(1.0 + v);
    if (log(1.0 - u) <= q) return scale * ret_val;
  }
  for (;;) {
    /* step 8: double-exponential sample */
    e = r_exp_rand();
    u = r_unif_rand();
    u = u + u - 1.0;
    if (u < 0.0) t = b - si * e;
    else t = b + si * e;
    /* step 9 */
    if (t >= -0.71874483771719) {
      v = t / (s + s);
      if (fabs(v) <= 0.25)
        q = q0 + 0.5 * t * t * ((((((a7 * v + a6) * v + a5) * v + a4) * v + a3) * v + a2) * v + a1) * v;
      else
        q = q0 - s * t + 0.25 * t * t + (s2 + s2) * log(1.0 + v);
      if (q > 0.0) {
        w = expm1(q);
        if (c * fabs(u) <= w * exp(e - 0.5 * t * t)) break;
      }
    }
  }
  x = s + 0.5 * t;
  return scale * x * x;
}

double r_rchisq(double df) {
  if (!isfinite(df) || df < 0.0) return NAN;
  return r_rgamma(df / 2.0, 2.0);
}

double r_rt1(double df) {
  if (isnan(df) || df <= 0.0) return NAN;
  if (!isfinite(df)) return r_norm_rand();
  double num = r_norm_rand();
  return num / sqrt(r_rchisq(df) / df);
}
void r_rt(int64_t n, double df, double* out) {
  for (int64_t i = 0; i < n; i++) out[i] = r_rt1(df);
}
void r_rgamma_vec(int64_t n, double a, double scale, double* out) {
  for (int64_t i = 0; i < n; i++) out[i] = r_rgamma(a, scale);
}
void r_exp_rand_vec(int64_t n, double* out) {
  for (int64_t i = 0; i < n; i++) out[i] = r_exp_rand();
}
