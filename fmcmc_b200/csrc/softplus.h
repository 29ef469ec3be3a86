// softplus.h — branch-free FP64 log(1 + exp(-a)), a >= 0, for the logistic family's epilogue.
//
// The logistic log-posterior (vignettes/workflow-with-fmcmc.Rmd:36-38) needs exactly one
// exp and one log1p per chain-step x observation; with CUDA's libm versions (slow paths,
// branches, ~110 non-FP64 instructions) they dominate the hot kernel (profiles/r01_v0_*).
// This version is straight-line code so four evaluations interleave in the FP64 pipe:
//   exp(-a) : n = rint(-a log2 e) (magic-number add), Cody-Waite reduction with fdlibm's
//             ln2_hi/ln2_lo, degree-11 near-minimax polynomial (Chebyshev interpolant,
//             rel. error 1.6e-17 on |r| <= ln2/2), scale by 2^n through the exponent bits;
//   log1p(e): u = 1 + e with its exact rounding error c carried along, u folded into
//             [sqrt(.5), sqrt(2)], s = f/(2+f) via reciprocal seed + one cubic Newton step,
//             fdlibm e_log.c's degree-14 odd polynomial (Lg1..Lg7, error < 2^-58.45).
// Host-compilable (plain C++) so tests/ can check it against mpmath on the CPU: max error
// 1 ulp-ish (see tests/test_softplus_cpu.py); the device differs only in the reciprocal seed.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define FM_HD __host__ __device__ __forceinline__
#else
#define FM_HD static inline
#endif

FM_HD double fm_rcp_seed(double x) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
#else
  return (double)(1.0f / (float)x);
#endif
}

FM_HD double fm_bits_to_double(int64_t b) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double(b);
#else
  double d;
  memcpy(&d, &b, 8);
  return d;
#endif
}
FM_HD int64_t fm_double_to_bits(double d) {
#ifdef __CUDA_ARCH__
  return __double_as_longlong(d);
#else
  int64_t b;
  memcpy(&b, &d, 8);
  return b;
#endif
}

// exp(-a) for a in [0, 708]
FM_HD double fm_exp_neg(double a) {
  const double x = -a;
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
  const double t = fma(x, 1.4426950408889634074, MAGIC);
  const double n = t - MAGIC;
  double r = fma(n, -6.93147180369123816490e-01, x);
  r = fma(n, -1.90821492927058770002e-10, r);
  double p = 2.5110049204818658e-08;
  p = fma(p, r, 2.763265472252779e-07);
  p = fma(p, r, 2.755724088722987e-06);
  p = fma(p, r, 2.4801485441561313e-05);
  p = fma(p, r, 0.00019841269890076403);
  p = fma(p, r, 0.0013888888952352863);
  p = fma(p, r, 0.008333333333319589);
  p = fma(p, r, 0.04166666666648795);
  p = fma(p, r, 0.1666666666666668);
  p = fma(p, r, 0.5000000000000019);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  // 2^n: the low word of t holds n (two's complement), n in [-1022, 0]
  const int64_t ni = (int64_t)(int32_t)(uint32_t)fm_double_to_bits(t);  // NaN in -> garbage scale, but p is NaN
  const double scale = fm_bits_to_double((ni + 1023) << 52);
  return p * scale;
}

FM_HD int32_t fm_hi_word(double d) { return (int32_t)(fm_double_to_bits(d) >> 32); }
FM_HD double fm_add_hi_word(double d, int32_t delta) {
  return fm_bits_to_double(fm_double_to_bits(d) + ((int64_t)delta << 32));
}

#ifndef FM_RCP_EXTRA_STEP
#define FM_RCP_EXTRA_STEP 0  // the reciprocal seed (>= 2^-20 accurate) + one cubic step is already < 2^-58
#endif

// log(1 + e) for e in [0, 1].  All selections are integer tests on the high word (ALU pipe), so the
// FP64 pipe only sees the arithmetic: 25 FP64 instructions.
FM_HD double fm_log1p_unit(double e) {
  const double u = 1.0 + e;
  const double c = e - (u - 1.0);  // exact: u + c == 1 + e
  // fold u in [1,2] into m in [sqrt(.5), sqrt(2)) like fdlibm e_log.c: test the high word against sqrt(2)
  const bool big = fm_hi_word(u) > 0x3FF6A09E;
  const double m = fm_add_hi_word(u, big ? -0x00100000 : 0);  // m = big ? u/2 : u (exact)
  const double f = m - 1.0;  // exact
  const double d = 2.0 + f;
  double rc = fm_rcp_seed(d);
  double er = fma(-d, rc, 1.0);
  rc = fma(rc, fma(er, er, er), rc);  // cubic step: rel. error ~ seed^3
#if FM_RCP_EXTRA_STEP
  er = fma(-d, rc, 1.0);
  rc = fma(rc, er, rc);
#endif
  const double s = f * rc;
  const double z = s * s;
  double R = 1.479819860511658591e-01;
  R = fma(R, z, 1.531383769920937332e-01);
  R = fma(R, z, 1.818357216161805012e-01);
  R = fma(R, z, 2.222219843214978396e-01);
  R = fma(R, z, 2.857142874366239149e-01);
  R = fma(R, z, 3.999999999940941908e-01);
  R = fma(R, z, 6.666666666666735130e-01);
  R = R * z;
  // log(m) = f - (hfsq - s (hfsq + R)), hfsq = f^2 / 2;  + [big] ln2 (hi/lo) + c / u  (c/u ~= c: |c| <= 2^-53)
  const double q = f * f;
  const double A = fma(0.5, q, R);
  const double lo = c + (big ? 1.90821492927058770002e-10 : 0.0);
  const double B = fma(s, A, lo);
  const double t = fma(0.5, q, -B);
  return (f - t) + (big ? 6.93147180369123816490e-01 : 0.0);
}

// log(1 + exp(-a)), a >= 0 (NaN is passed through: the clamp is an integer test that NaN fails)
FM_HD double fm_softplus_neg(double a) {
  const bool over = fm_hi_word(a) >= 0x40862000 && fm_hi_word(a) < 0x7FF00000;  // a >= 708 and finite/inf
  a = (fm_hi_word(a) >= 0x7FF00000 && !(a != a)) ? 708.0 : a;                  // +inf
  a = over ? 708.0 : a;
  return fm_log1p_unit(fm_exp_neg(a));
}

// ---- table-driven log(1 + exp(-a)) for the observation-tiled hot kernel ---------------------
// a = k/32 + d, |d| <= 1/64.  With E = exp(-k/32), S = E/(1+E), G = log1p(E) (table, 16 B/entry):
//   log1p(exp(-a)) = G + log1p(S * expm1(-d)).
// |S expm1(-d)| <= 0.0079 and the correction is <= 1.6 % of the result, so two degree-5 polynomials
// (near-minimax, Chebyshev-node interpolants; rel. error 9e-17 and 1.1e-15 on their intervals) keep the
// total below ~1.5 ulp: 16 FP64 instructions + one LDS.128, against 41 for the table-free version above.
// Arguments above FM_SP_AMAX are clamped (the result is then < 2^-92 in absolute terms, far below one
// ulp of any log-likelihood term it is added to).  NaN is NOT propagated here: callers test eta's bits.
#define FM_SP_H 32
#define FM_SP_AMAX 64
#define FM_SP_ENTRIES (FM_SP_AMAX * FM_SP_H + 1)

// entry k: tab[2k] = S_k, tab[2k+1] = G_k; computed in long double and rounded once
static inline void fm_softplus_table_fill(double* tab) {
  for (int k = 0; k < FM_SP_ENTRIES; k++) {
    const long double x = (long double)k / FM_SP_H;
    const long double E = expl(-x);
    tab[2 * k] = (double)(E / (1.0L + E));
    tab[2 * k + 1] = (double)log1pl(E);
  }
}

// a >= 0 with a <= FM_SP_AMAX already enforced by the caller; k = index, d = a - k/32
FM_HD double fm_softplus_tab_core(double d, double S, double G) {
  double q = -0x1.6c175d75f692ap-10;
  q = fma(q, d, 0x1.1111ad1af8af9p-7);
  q = fma(q, d, -0x1.5555555538138p-5);
  q = fma(q, d, 0x1.555555551ad1ap-3);
  q = fma(q, d, -0x1.0000000000000p-1);
  q = fma(q, d, 0x1.0000000000000p+0);   // q = expm1(-d) / (-d)
  const double v = (S * d) * -q;          // S * expm1(-d)
  double L = -0x1.555b6df3e4efdp-3;
  L = fma(L, v, 0x1.99a091298881fp-3);
  L = fma(L, v, -0x1.fffffff6b5a52p-3);
  L = fma(L, v, 0x1.555555500646bp-2);
  L = fma(L, v, -0x1.0000000000008p-1);
  L = fma(L, v, 0x1.0000000000005p+0);   // L = log1p(v) / v
  return fma(v, L, G);
}

// reference composition (host tests; the device kernel inlines the same steps around its LDS)
FM_HD double fm_softplus_tab(double a, const double* tab) {
  a = (fm_hi_word(a) >= 0x40500000) ? (double)FM_SP_AMAX : a;  // >= 64, +inf, NaN -> 64
  const double MAGIC = 6755399441055744.0;
  const double t = fma(a, (double)FM_SP_H, MAGIC);
  const int32_t k = (int32_t)(uint32_t)fm_double_to_bits(t);
  const double d = fma(t - MAGIC, -1.0 / FM_SP_H, a);
  return fm_softplus_tab_core(d, tab[2 * k], tab[2 * k + 1]);
}

// ---- finer table for the split-integer kernel (tiled_i8.cuh), where every FP64 instruction of the epilogue counts ----
// a = k/128 + d, |d| <= 1/256, a clamped to 40 (log1p(exp(-40)) = 4.2e-18, below half an ulp of any sum it enters):
// 5121 entries = 80 KB of shared memory, and plain degree-4 Taylor polynomials are exact enough
//   expm1(-d) / (-d) = 1 - d/2 + d^2/6 - d^3/24 + d^4/120      (next term d^5/720 <= 1.3e-15)
//   log1p(v) / v     = 1 - v/2 + v^2/3 - v^3/4 + v^4/5          (|v| <= 2e-3: next term 5e-15, times |v| in the result)
// 13 FP64 instructions + one LDS.128 against 15 for the 32-per-unit table.
#define FM_SP4_H 128
#define FM_SP4_AMAX 40
#define FM_SP4_ENTRIES (FM_SP4_AMAX * FM_SP4_H + 1)
static inline void fm_softplus_table4_fill(double* tab) {
  for (int k = 0; k < FM_SP4_ENTRIES; k++) {
    const long double x = (long double)k / FM_SP4_H;
    const long double E = expl(-x);
    tab[2 * k] = (double)(E / (1.0L + E));
    tab[2 * k + 1] = (double)log1pl(E);
  }
}
FM_HD double fm_softplus_tab4_core(double d, double S, double G) {
  double q = 1.0 / 120.0;
  q = fma(q, d, -1.0 / 24.0);
  q = fma(q, d, 1.0 / 6.0);
  q = fma(q, d, -0.5);
  q = fma(q, d, 1.0);
  const double v = (S * d) * -q;
  double L = 0.2;
  L = fma(L, v, -0.25);
  L = fma(L, v, 1.0 / 3.0);
  L = fma(L, v, -0.5);
  L = fma(L, v, 1.0);
  return fma(v, L, G);
}
// reference composition (host tests)
FM_HD double fm_softplus_tab4(double a, const double* tab) {
  a = (fm_hi_word(a) >= 0x40440000) ? (double)FM_SP4_AMAX : a;  // >= 40, +inf, NaN -> 40
  const double MAGIC = 6755399441055744.0;
  const double t = fma(a, (double)FM_SP4_H, MAGIC);
  const int32_t k = (int32_t)(uint32_t)fm_double_to_bits(t);
  const double d = fma(t - MAGIC, -1.0 / FM_SP4_H, a);
  return fm_softplus_tab4_core(d, tab[2 * k], tab[2 * k + 1]);
}

// ---- 256-per-unit table: 10 241 entries = 160 KB of shared memory, |d| <= 1/512, CUBIC near-minimax polynomials
// (Chebyshev-node interpolants; |error| 1.5e-14 and 2.5e-14 relative on q and L, i.e. <= 4e-17 absolute in the result):
// 11 FP64 instructions + one LDS.128.
#define FM_SP8_H 256
#define FM_SP8_AMAX 40
#define FM_SP8_ENTRIES (FM_SP8_AMAX * FM_SP8_H + 1)
static inline void fm_softplus_table8_fill(double* tab) {
  for (int k = 0; k < FM_SP8_ENTRIES; k++) {
    const long double x = (long double)k / FM_SP8_H;
    const long double E = expl(-x);
    tab[2 * k] = (double)(E / (1.0L + E));
    tab[2 * k + 1] = (double)log1pl(E);
  }
}
FM_HD double fm_softplus_tab8_core(double d, double S, double G) {
  double q = -0x1.5555582d82db0p-5;
  q = fma(q, d, 0x1.55555999999f5p-3);
  q = fma(q, d, -0x1.fffffffffffd2p-2);
  q = fma(q, d, 0x1.fffffffffff77p-1);    // q = expm1(-d) / (-d), |d| <= 1/512
  const double v = (S * d) * -q;
  double L = -0x1.00000b2f503bap-2;
  L = fma(L, v, 0x1.555562c14f2f7p-2);
  L = fma(L, v, -0x1.ffffffffffe89p-2);
  L = fma(L, v, 0x1.fffffffffff1fp-1);    // L = log1p(v) / v, |v| <= 1e-3
  return fma(v, L, G);
}
FM_HD double fm_softplus_tab8(double a, const double* tab) {
  a = (fm_hi_word(a) >= 0x40440000) ? (double)FM_SP8_AMAX : a;  // >= 40, +inf, NaN -> 40
  const double MAGIC = 6755399441055744.0;
  const double t = fma(a, (double)FM_SP8_H, MAGIC);
  const int32_t k = (int32_t)(uint32_t)fm_double_to_bits(t);
  const double d = fma(t - MAGIC, -1.0 / FM_SP8_H, a);
  return fm_softplus_tab8_core(d, tab[2 * k], tab[2 * k + 1]);
}

// ---- log(2 cosh(a / 2)) = a / 2 + log(1 + exp(-a)) on the same 256-per-unit grid (the binary-logistic epilogue of
// tiled_i8.cuh accumulates exactly this even function of eta: tiled_i8.cuh, i8_logistic_lcosh).  h is analytic and all
// its derivatives are polynomials in tau = h' = tanh(a/2) / 2:  with u = 1/4 - tau^2
//   h'' = u,  h''' = -2 tau u,  h'''' = u (1/2 - 6 u) ... so the degree-4 Taylor expansion around the table point needs
// only (tau_k, T_k = h(k/256)) from the table, 16 B per entry as before:
//   h(k/256 + d) = T + d (tau + d u (1/2 + d (-tau/3 + d (1/24 - u/4)))),  remainder <= 0.13 d^5 / 120 = 3e-17 (|d| <= 1/512)
// 9 FP64 instructions including the accumulation (against 11 for G + log1p(S expm1(-d)) plus one for sum |eta| / 2).
// The last entry (a = 40) has tau = 1/2 exactly in double, hence u = 0 and h = T + d / 2 there.  Arguments beyond the table
// are the caller's business: the kernel either proves |eta| <= 39.9 for all its chains (no clamp at all) or clamps the
// argument and adds the excess |eta| / 2 apart (tiled_i8.cuh).
static inline void fm_lcosh_table8_fill(double* tab) {
  for (int k = 0; k < FM_SP8_ENTRIES; k++) {
    const long double x = (long double)k / FM_SP8_H;
    const long double E = expl(-x);
    tab[2 * k] = (double)(0.5L - E / (1.0L + E));          // tau = 1/2 - sigma(-x)
    tab[2 * k + 1] = (double)(0.5L * x + log1pl(E));       // T
  }
}
FM_HD double fm_lcosh_tab8_core(double d, double tau, double T) {
  const double u = fma(-tau, tau, 0.25);
  const double i1 = fma(u, -0.25, 1.0 / 24.0);
  const double i2 = fma(d, i1, tau * (-1.0 / 3.0));
  const double i3 = fma(d, i2, 0.5);
  const double P = u * i3;
  const double Q = fma(d, P, tau);
  return fma(d, Q, T);
}
// ---- the degree-3 variant (the hot loop of the split-integer kernel when every chain of a CTA stays inside the table) ----
// Dropping the d^4 term u (1/24 - u/4) d^4 saves two FP64 instructions of nine.  What it leaves out is at most 7.6e-14 (a = 0,
// |d| = 1/512), EVEN in d, and the remainder of a sum over many observations is what matters: its mean over the uniformly
// distributed remainder d, c4 delta^4 / 5, is folded into the table's T (fm_lcosh_table8m_fill), so the error per evaluation
// is zero-mean with rms <= 2e-14 - fifty times below the kernel's own slicing error in eta (~1e-12 tau per evaluation) - and
// averages out like it: ~3e-17 relative in a log-posterior over 1e6 observations (tests/test_softplus_cpu.py).
static inline void fm_lcosh_table8m_fill(double* tab) {
  const long double delta4_5 = powl(0.5L / FM_SP8_H, 4) / 5.0L;
  for (int k = 0; k < FM_SP8_ENTRIES; k++) {
    const long double x = (long double)k / FM_SP8_H;
    const long double E = expl(-x);
    const long double tau = 0.5L - E / (1.0L + E);
    const long double u = 0.25L - tau * tau;
    tab[2 * k] = (double)tau;
    tab[2 * k + 1] = (double)(0.5L * x + log1pl(E) + u * (1.0L / 24.0L - u / 4.0L) * delta4_5);
  }
}
FM_HD double fm_lcosh_tab8m_core(double d, double tau, double T) {
  const double u = fma(-tau, tau, 0.25);
  const double i3 = fma(d, tau * (-1.0 / 3.0), 0.5);
  const double P = u * i3;
  const double Q = fma(d, P, tau);
  return fma(d, Q, T);
}
FM_HD double fm_lcosh_tab8m(double a, const double* tab) {   // host reference: 0 <= a < 40
  const double MAGICH = 26388279066624.0;
  const double t2 = a + MAGICH;
  uint32_t k = (uint32_t)fm_double_to_bits(t2);
  k = k > (uint32_t)(FM_SP8_ENTRIES - 1) ? (uint32_t)(FM_SP8_ENTRIES - 1) : k;
  const double d = a + (MAGICH - t2);
  return fm_lcosh_tab8m_core(d, tab[2 * k], tab[2 * k + 1]);
}
// ---- the bank-group-replicated cubic table: two conflict-free loads instead of three FP64 instructions ----
// The hot loop's time is its issued instructions (profiles/r01_i8_findings.md, section 4b), and three of the ten FP64
// instructions of the degree-3 core only rebuild the quadratic and cubic coefficients from tau (u = 1/4 - tau^2, -tau/3, their
// product).  Reading them from the table instead costs a second LDS.128 - which the shared-memory data pipe (128 B per clock
// and SM) only affords if the gather is free of bank conflicts: an LDS.128 is served a quarter-warp (8 lanes) at a time, 8
// random 16-byte entries fall on the 8 four-bank groups with a maximum load of ~2.6, ~10 wavefronts instead of 4.  Here every
// table point k is 256 bytes: eight copies of (c1, c0) followed by eight copies of (c2, c3); lane l reads copy l % 8 of each,
// i.e. bank group l % 8 whatever k is - each load is exactly 4 wavefronts.  184 KB of shared memory hold 721 points: a
// 64-per-unit grid up to a = 11.25; chains whose bound on |eta| is larger use the 256-per-unit tables above.
// With four free coefficients per cell the cubic is the LEAST-SQUARES fit of h on the cell (uniform weight; Legendre projection
// of the degree-6 Taylor expansion: d^4 -> 6/7 D^2 d^2 - 3/35 D^4, d^5 -> 10/9 D^2 d^3 - 5/21 D^4 d, d^6 -> 5/7 D^4 d^2 - 2/21 D^6,
// D = 1/128 the half-width), so the error is orthogonal to 1, d, d^2, d^3 over the cell - zero-mean in particular - and equals
// (8/35) c4 D^4 P4(d / D): <= 4.5e-12 at the cell edges for a = 0, rms <= 1.5e-12, decaying like sech^2(a / 2); that is the size
// of the slicing error the same evaluation already carries in eta (5.5e-12 of the largest term) and averages out like it.
#define FM_LC6_H 64
#define FM_LC6_REP 8
#define FM_LC6_POINT_BYTES (2 * 16 * FM_LC6_REP)
#define FM_LC6_ENTRIES_MAX 721
static inline void fm_lcosh_cubic_coeffs(long double x, long double D, long double c[4]) {   // LSQ cubic of h on [x - D, x + D]
  const long double E = expl(-x);
  const long double tau = 0.5L - E / (1.0L + E), u = 0.25L - tau * tau;
  const long double c0 = 0.5L * x + log1pl(E), c1 = tau, c2 = u / 2.0L, c3 = -tau * u / 3.0L;
  const long double c4 = u * (1.0L - 6.0L * u) / 24.0L, c5 = -2.0L * tau * u * (1.0L - 12.0L * u) / 120.0L;
  const long double c6 = -2.0L * (u * u * (1.0L - 12.0L * u) - 2.0L * tau * tau * u * (1.0L - 24.0L * u)) / 720.0L;
  const long double D2 = D * D, D4 = D2 * D2, D6 = D4 * D2;
  c[0] = c0 - (3.0L / 35.0L) * c4 * D4 - (2.0L / 21.0L) * c6 * D6;
  c[1] = c1 - (5.0L / 21.0L) * c5 * D4;
  c[2] = c2 + (6.0L / 7.0L) * c4 * D2 + (5.0L / 7.0L) * c6 * D4;
  c[3] = c3 + (10.0L / 9.0L) * c5 * D2;
}
static inline void fm_lcosh_table6r_fill(double* tab) {   // tab: FM_LC6_ENTRIES_MAX * FM_LC6_POINT_BYTES / 8 doubles
  for (int k = 0; k < FM_LC6_ENTRIES_MAX; k++) {
    long double c[4];
    fm_lcosh_cubic_coeffs((long double)k / FM_LC6_H, 0.5L / FM_LC6_H, c);
    double* pt = tab + (size_t)k * (FM_LC6_POINT_BYTES / 8);
    for (int r = 0; r < FM_LC6_REP; r++) {
      pt[2 * r] = (double)c[1];
      pt[2 * r + 1] = (double)c[0];
      pt[2 * FM_LC6_REP + 2 * r] = (double)c[2];
      pt[2 * FM_LC6_REP + 2 * r + 1] = (double)c[3];
    }
  }
}
FM_HD double fm_lcosh_tab6r_core(double d, double c1, double c0, double c2, double c3) {
  const double q1 = fma(d, c3, c2);
  const double q2 = fma(d, q1, c1);
  return fma(d, q2, c0);
}
FM_HD double fm_lcosh_tab6r(double a, const double* tab, int r) {   // host reference: 0 <= a <= 11.25, copy r
  const double MAGICH = 105553116266496.0;  // 1.5 * 2^46: ulp = 1/64
  const double t2 = a + MAGICH;
  uint32_t k = (uint32_t)fm_double_to_bits(t2);
  k = k > (uint32_t)(FM_LC6_ENTRIES_MAX - 1) ? (uint32_t)(FM_LC6_ENTRIES_MAX - 1) : k;
  const double d = a + (MAGICH - t2);
  const double* pt = tab + (size_t)k * (FM_LC6_POINT_BYTES / 8) + 2 * (size_t)(r & (FM_LC6_REP - 1));
  return fm_lcosh_tab6r_core(d, pt[0], pt[1], pt[2 * FM_LC6_REP], pt[2 * FM_LC6_REP + 1]);
}
// reference composition (host tests): 0 <= a <= 40
FM_HD double fm_lcosh_tab8(double a, const double* tab) {
  const double MAGICH = 26388279066624.0;  // 1.5 * 2^44: ulp = 1/256
  const double t2 = a + MAGICH;
  uint32_t k = (uint32_t)fm_double_to_bits(t2);
  k = k > (uint32_t)(FM_SP8_ENTRIES - 1) ? (uint32_t)(FM_SP8_ENTRIES - 1) : k;
  const double d = a + (MAGICH - t2);
  return fm_lcosh_tab8_core(d, tab[2 * k], tab[2 * k + 1]);
}
