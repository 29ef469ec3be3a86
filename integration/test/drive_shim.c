/*
 * drive_shim.c - plays the part of R/device.R's MCMC_device() in C: builds the lists the R glue passes to .Call and calls
 * the shim's entry points (integration/src/shim.c) against the R C-API stand-in (integration/rstub/).  Test infrastructure:
 * tests/test_shim_compile.py compiles it, runs it on the GPU box and compares what comes back with the same runs made
 * through the Python mirror (bit for bit: same library, same Philox streams).
 *
 *   drive_shim <input.bin> <output.bin>     two scenarios on the README model (README.md:112-139):
 *        A  kernel_normal_reflective, 3 bulks of `bulk` rows with FMCMC_RUN_APPEND, then C_fmcmc_gelman  (R/mcmc.R:901-968)
 *        B  kernel_adapt(warmup = 20), 2 bulks, kernel state carried through state$istate / state$dstate (R/mcmc.R:629-631)
 *        then the error path (burnin >= nsteps -> Rf_error with the reference's message) and the finalizer
 *   drive_shim --no-gpu                     marshalling-only check for machines without a GPU: C_fmcmc_model_create must
 *                                           fail through Rf_error with the library's message, PROTECT stack balanced
 *
 * input.bin : int32 n, nchains, bulk, k(=3); double X[n], y[n], initial[nchains][k]
 * output.bin: scenario A: for each of 3 bulks ans[rows*k*nchains] (R layout [row, param, chain]), logpost[rows*nchains];
 *             psrf[k], mpsrf, niter;   scenario B: for each of 2 bulks ans, then istate[4*nchains], dstate[dlen*nchains]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#include "fmcmc_b200.h"

SEXP C_fmcmc_model_create(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP C_fmcmc_model_free(SEXP);
SEXP C_fmcmc_store_reset(SEXP, SEXP, SEXP, SEXP);
SEXP C_fmcmc_run(SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP C_fmcmc_gelman(SEXP, SEXP, SEXP, SEXP);
SEXP C_fmcmc_reflect(SEXP, SEXP, SEXP, SEXP, SEXP);
void R_init_fmcmc(DllInfo*);

#define CHECK_BALANCED(what)                                                                   \
  do {                                                                                         \
    rstub_end_call();                                                                          \
    if (rstub_protect_depth() != 0) {                                                          \
      fprintf(stderr, "%s: PROTECT stack imbalance (%d)\n", what, rstub_protect_depth());      \
      return 4;                                                                                \
    }                                                                                          \
  } while (0)

static SEXP named(int n, const char** names, SEXP* values) { return rstub_named_list(n, names, values); }
static SEXP get(SEXP lst, const char* name) {
  SEXP nm = Rf_getAttrib(lst, R_NamesSymbol);
  for (R_xlen_t i = 0; i < XLENGTH(lst); i++)
    if (!strcmp(CHAR(STRING_ELT(nm, i)), name)) return VECTOR_ELT(lst, i);
  return R_NilValue;
}

static SEXP run_list(double nsteps, double burnin, int nchains, int flags, SEXP initial) {
  const char* nm[] = {"nsteps", "burnin", "thin", "nchains", "flags", "chain_offset", "nchains_total", "initial"};
  SEXP v[] = {ScalarReal(nsteps), ScalarReal(burnin), ScalarReal(1), ScalarInteger(nchains), ScalarInteger(flags),
              ScalarReal(0), ScalarReal(0), initial};
  return named(8, nm, v);
}
static SEXP stream_list(double seed, double run_index) {
  const char* nm[] = {"mode", "seed", "run_index", "kdraw", "logu", "z"};
  SEXP v[] = {ScalarInteger(FMCMC_STREAM_PHILOX), ScalarReal(seed), ScalarReal(run_index), ScalarInteger(0), R_NilValue, R_NilValue};
  return named(6, nm, v);
}
/* what kernel_to_spec() of R/device.R extracts from the kernel environment */
static SEXP kernel_list(int type, int k, double scale, double warmup) {
  double mu[3] = {0, 0, 0}, sc[3] = {scale, scale, scale}, mn[3] = {-1, -1, -1}, mx[3] = {1, 1, 1};
  double lb[3] = {-5.0, 0.0, 0.0}, ub[3] = {5.0, 5.0, 5.0};
  unsigned char fixed[3] = {0, 0, 0};
  const char* nm[] = {"type", "k", "scheme", "order", "mu", "scale", "min", "max", "lb", "ub", "fixed", "warmup", "freq", "bw",
                      "until", "eps", "Sd", "arate", "nadapt", "constr", "mvn_method"};
  SEXP v[] = {ScalarInteger(type), ScalarInteger(k), ScalarInteger(FMCMC_SCHEME_JOINT), R_NilValue, rstub_real(k, mu), rstub_real(k, sc),
              rstub_real(k, mn), rstub_real(k, mx), rstub_real(k, lb), rstub_real(k, ub), rstub_raw(k, fixed), ScalarReal(warmup),
              ScalarReal(1), ScalarReal(0), ScalarReal(1.0 / 0.0), ScalarReal(1e-4), ScalarReal(-1), ScalarReal(0.234), R_NilValue,
              R_NilValue, ScalarInteger(FMCMC_MVN_CHOLESKY)};
  return named(21, nm, v);
}
static SEXP state_list(SEXP ist, SEXP dst) {
  const char* nm[] = {"istate", "dstate"};
  SEXP v[] = {ist, dst};
  return named(2, nm, v);
}
static void put(FILE* f, SEXP x) { fwrite(REAL(x), sizeof(double), (size_t)XLENGTH(x), f); }

int main(int argc, char** argv) {
  DllInfo dll;
  memset(&dll, 0, sizeof dll);
  R_init_fmcmc(&dll); /* registration table: every routine the R code .Call()s must be there with its arity */
  int nreg = 0;
  for (const R_CallMethodDef* c = dll.calls; c && c->name; c++) nreg++;
  if (nreg != 7 || dll.dynamic_symbols != FALSE) {
    fprintf(stderr, "registration: %d routines\n", nreg);
    return 4;
  }
  if (argc == 2 && !strcmp(argv[1], "--no-gpu")) {
    const double X[4] = {0.1, -0.2, 0.3, 0.4}, y[4] = {1, 2, 3, 4}, hyper[4] = {0, 0, 0, 0};
    SEXP Xm = Rf_allocMatrix(REALSXP, 4, 1);
    memcpy(REAL(Xm), X, sizeof X);
    rstub_error_armed = 1;
    if (setjmp(rstub_error_jmp) == 0) {
      /* device 99 does not exist anywhere: the library's error must come back through Rf_error */
      (void)C_fmcmc_model_create(ScalarInteger(FMCMC_FAMILY_GAUSSIAN_LM), ScalarInteger(3), Xm, rstub_real(4, y), R_NilValue,
                                 ScalarInteger(0), rstub_real(4, hyper), ScalarInteger(99));
      fprintf(stderr, "C_fmcmc_model_create on device 99 did not fail\n");
      return 5;
    }
    CHECK_BALANCED("model_create error path");
    /* a mismatch between X and y is caught by the shim itself */
    if (setjmp(rstub_error_jmp) == 0) {
      (void)C_fmcmc_model_create(ScalarInteger(FMCMC_FAMILY_GAUSSIAN_LM), ScalarInteger(3), Xm, rstub_real(3, y), R_NilValue,
                                 ScalarInteger(0), rstub_real(4, hyper), ScalarInteger(0));
      return 5;
    }
    if (!strstr(rstub_error_msg, "rows")) return 6;
    printf("no-gpu: errors propagate through Rf_error (\"%s\"), PROTECT stack balanced, %d routines registered\n",
           rstub_error_msg, nreg);
    return 0;
  }
  if (argc != 3) {
    fprintf(stderr, "usage: drive_shim <input.bin> <output.bin> | --no-gpu\n");
    return 2;
  }
  FILE* in = fopen(argv[1], "rb");
  if (!in) return 2;
  int hdr[4];
  if (fread(hdr, sizeof(int), 4, in) != 4) return 2;
  const int n = hdr[0], C = hdr[1], bulk = hdr[2], k = hdr[3];
  SEXP X = Rf_allocMatrix(REALSXP, n, 1), y = Rf_allocVector(REALSXP, n), init = Rf_allocMatrix(REALSXP, k, C); /* t(initial) */
  if (fread(REAL(X), 8, (size_t)n, in) != (size_t)n || fread(REAL(y), 8, (size_t)n, in) != (size_t)n ||
      fread(REAL(init), 8, (size_t)C * k, in) != (size_t)C * k)
    return 2;
  fclose(in);
  FILE* out = fopen(argv[2], "wb");
  if (!out) return 2;
  const double hyper[4] = {0, 0, 0, 0};

  SEXP model = C_fmcmc_model_create(ScalarInteger(FMCMC_FAMILY_GAUSSIAN_LM), ScalarInteger(FMCMC_MODEL_INTERCEPT | FMCMC_MODEL_GUARD), X, y,
                                    R_NilValue, ScalarInteger(0), rstub_real(4, hyper), ScalarInteger(0));
  CHECK_BALANCED("model_create");

  /* ---- scenario A: the bulk loop with the device Gelman check ---- */
  C_fmcmc_store_reset(model, ScalarInteger(C), ScalarInteger(k), ScalarReal(3.0 * bulk));
  SEXP spec = kernel_list(FMCMC_KERNEL_NORMAL_REFLECTIVE, k, 0.05, 0);
  SEXP state = state_list(R_NilValue, R_NilValue); /* fresh kernel: all zeros */
  for (int b = 0; b < 3; b++) {
    SEXP res = C_fmcmc_run(model, run_list(bulk, 0, C, FMCMC_RUN_APPEND, b == 0 ? init : R_NilValue), spec, state, stream_list(11, b));
    CHECK_BALANCED("run A");
    SEXP ans = get(res, "ans");
    SEXP dim = Rf_getAttrib(ans, R_DimSymbol);
    if (INTEGER(dim)[0] != bulk || INTEGER(dim)[1] != k || INTEGER(dim)[2] != C) return 7;
    if (REAL(get(res, "report"))[0] != bulk) return 7;
    put(out, ans);
    put(out, get(res, "logpost"));
    state = state_list(get(res, "istate"), get(res, "dstate"));
  }
  unsigned char free_mask[3] = {1, 1, 1};
  SEXP g = C_fmcmc_gelman(model, rstub_raw(k, free_mask), ScalarReal(1), ScalarReal(1));
  CHECK_BALANCED("gelman");
  put(out, VECTOR_ELT(g, 0));
  put(out, VECTOR_ELT(g, 1));
  put(out, VECTOR_ELT(g, 2));

  /* ---- scenario B: an adaptive kernel whose state travels through R objects between the calls ---- */
  spec = kernel_list(FMCMC_KERNEL_ADAPT, k, 1.0, 20);
  state = state_list(R_NilValue, R_NilValue);
  for (int b = 0; b < 2; b++) {
    SEXP res = C_fmcmc_run(model, run_list(bulk, 0, C, 0, b == 0 ? init : R_NilValue), spec, state, stream_list(12, b));
    CHECK_BALANCED("run B");
    put(out, get(res, "ans"));
    state = state_list(get(res, "istate"), get(res, "dstate"));
  }
  put(out, get(state, "istate"));
  put(out, get(state, "dstate"));
  if (REAL(get(state, "istate"))[0] != 2.0 * (bulk - 1)) { /* abs_iter: one proposal per row after the first, two bulks */
    fprintf(stderr, "abs_iter = %g\n", REAL(get(state, "istate"))[0]);
    return 8;
  }

  /* ---- exported helper: reflect_on_boundaries(c(2.3, -0.4), lb = 0, ub = 1.5) = c(0.7, 0.4) ---- */
  const double xs[2] = {2.3, -0.4}, lb1[1] = {0.0}, ub1[1] = {1.5};
  SEXP rf = C_fmcmc_reflect(rstub_real(2, xs), rstub_real(1, lb1), rstub_real(1, ub1), R_NilValue, ScalarInteger(0));
  CHECK_BALANCED("reflect");
  put(out, rf);

  /* ---- error path: the reference's message comes back through Rf_error, nothing leaks, the model stays usable ---- */
  rstub_error_armed = 1;
  if (setjmp(rstub_error_jmp) == 0) {
    (void)C_fmcmc_run(model, run_list(10, 10, C, 0, init), spec, state_list(R_NilValue, R_NilValue), stream_list(1, 0));
    fprintf(stderr, "burnin >= nsteps did not fail\n");
    return 5;
  }
  CHECK_BALANCED("run error path");
  if (!strstr(rstub_error_msg, "burnin")) {
    fprintf(stderr, "unexpected message: %s\n", rstub_error_msg);
    return 6;
  }
  printf("error path: \"%s\"\n", rstub_error_msg);
  rstub_error_armed = 0;
  (void)C_fmcmc_run(model, run_list(5, 0, C, 0, init), spec, state_list(R_NilValue, R_NilValue), stream_list(1, 0));
  CHECK_BALANCED("run after error");

  /* ---- garbage collection at exit: the finalizer frees the device model exactly once ---- */
  rstub_run_finalizers();
  if (R_ExternalPtrAddr(model) != NULL) return 9;
  C_fmcmc_model_free(model); /* a second, explicit free is a no-op */
  fclose(out);
  printf("ok: %d chains x %d rows, k = %d\n", C, bulk, k);
  return 0;
}
