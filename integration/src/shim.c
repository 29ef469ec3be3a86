/*
 * shim.c - the .Call layer between the R package fmcmc and libfmcmcb200.so (goes to src/shim.c of the reference package).
 *
 * Thin by design: no logic, only marshalling between R objects and the PODs of include/fmcmc_b200.h.  Every entry point
 * names the reference code whose work it hands to the device:
 *   C_fmcmc_model_create   the user closure `fun` + its captured data  (README.md:128-139, vignettes/workflow-with-fmcmc.Rmd:35-41)
 *   C_fmcmc_run            MCMC_without_conv_checker                   (R/mcmc.R:485-838)
 *   C_fmcmc_store_reset    the accumulating `ans` of the bulk loop     (R/mcmc.R:947)
 *   C_fmcmc_gelman         conv_checker(ans) -> coda::gelman.diag      (R/mcmc.R:968, R/convergence.R:207)
 *   C_fmcmc_cov_recursive  cov_recursive / mean_recursive              (R/recursive.R:63-139)
 *   C_fmcmc_reflect        reflect_on_boundaries                       (R/kernel.R:450-493)
 *
 * R is absent from the build image, so this file is compiled in the test-suite against integration/rstub/ (a minimal
 * stand-in for R's C API) with -Wall -Wextra -Werror, and driven from C by integration/test/drive_shim.c
 * (tests/test_shim_compile.py).  With real R it is compiled against R's own headers; nothing here depends on the stub.
 *
 * Threading / interrupts: every .Call comes from the R main thread; the library never calls back into R; Rf_error is raised
 * only after the library call has returned (its scratch is owned by the handle, so the longjmp leaks nothing);
 * R_CheckUserInterrupt() belongs between bulks in the R loop, not inside a call.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "fmcmc_b200.h"

static void model_finalizer(SEXP ptr) { /* device memory is owned by the handle */
  fmcmc_model* m = (fmcmc_model*)R_ExternalPtrAddr(ptr);
  if (m) {
    fmcmc_model_free(m);
    R_ClearExternalPtr(ptr);
  }
}
static fmcmc_model* model_of(SEXP ptr) {
  fmcmc_model* m = (fmcmc_model*)R_ExternalPtrAddr(ptr);
  if (!m) error("the fmcmc device model was freed");
  return m;
}
static SEXP elt(SEXP lst, const char* name) { /* list element by name, or R_NilValue */
  SEXP nm = getAttrib(lst, R_NamesSymbol);
  if (isNull(nm)) return R_NilValue;
  for (R_xlen_t i = 0; i < XLENGTH(lst); i++)
    if (!strcmp(CHAR(STRING_ELT(nm, i)), name)) return VECTOR_ELT(lst, i);
  return R_NilValue;
}
static const double* dbl_or_null(SEXP x) { return isNull(x) ? NULL : REAL(x); }
static double real_or(SEXP x, double dflt) { return isNull(x) ? dflt : asReal(x); }
static int int_or(SEXP x, int dflt) { return isNull(x) ? dflt : asInteger(x); }

/* .Call(C_fmcmc_model_create, family, flags, X, y, group, n_groups, hyper, device) -> external pointer */
SEXP C_fmcmc_model_create(SEXP family, SEXP flags, SEXP X, SEXP y, SEXP group, SEXP n_groups, SEXP hyper, SEXP device) {
  fmcmc_model_desc d;
  memset(&d, 0, sizeof d);
  d.family = asInteger(family);
  d.flags = (uint32_t)asInteger(flags);
  d.n = (int64_t)XLENGTH(y);
  d.p_x = isNull(X) ? 0 : ncols(X);
  if (!isNull(X) && (int64_t)nrows(X) != d.n) error("X has %d rows but y has %lld elements", nrows(X), (long long)d.n);
  d.X = dbl_or_null(X); /* R matrices are column-major: passed as they are */
  d.y = REAL(y);
  d.group = isNull(group) ? NULL : INTEGER(group); /* 0-based, prepared by the R glue */
  d.n_groups = int_or(n_groups, 0);
  for (int i = 0; i < 4 && i < LENGTH(hyper); i++) d.hyper[i] = REAL(hyper)[i];
  char err[1024] = "";
  fmcmc_model* m = NULL;
  const int rc = fmcmc_model_create(&d, asInteger(device), &m, err, sizeof err);
  if (rc) error("%s", err); /* nothing is allocated on failure */
  SEXP ptr = PROTECT(R_MakeExternalPtr(m, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, model_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}

/* .Call(C_fmcmc_model_free, model): release the HBM copy now instead of at garbage collection */
SEXP C_fmcmc_model_free(SEXP model) {
  model_finalizer(model);
  return R_NilValue;
}

/* .Call(C_fmcmc_store_reset, model, nchains, k, capacity_rows): room for the samples the bulk loop accumulates */
SEXP C_fmcmc_store_reset(SEXP model, SEXP nchains, SEXP k, SEXP capacity_rows) {
  char err[1024] = "";
  if (fmcmc_store_reset(model_of(model), asInteger(nchains), asInteger(k), (int64_t)asReal(capacity_rows), err, sizeof err))
    error("%s", err);
  return R_NilValue;
}

/* .Call(C_fmcmc_run, model, run, kernel, state, stream)
 * run    = list(nsteps, burnin, thin, nchains, flags, chain_offset, nchains_total, initial = t(initial) or NULL)
 * kernel = list(type, k, scheme, order, mu, scale, min, max, lb, ub, fixed (raw), warmup, freq, bw, until, eps, Sd, arate,
 *               nadapt, constr, mvn_method)   -- vectors already recycled / NA-processed (R/kernel.R:1-41)
 * state  = list(istate = numeric 4 x nchains (abs_iter, flags, nerrors, scratch as plain numbers), dstate = numeric dlen x nchains)
 * stream = list(mode, seed, run_index, kdraw, logu, z)
 * returns list(ans, draws, logpost, report, istate, dstate); ans / draws are [rows, k, chain] arrays (FMCMC_RUN_COLMAJOR),
 * so ans[, , c] is chain c's matrix without a transpose. */
SEXP C_fmcmc_run(SEXP model, SEXP run, SEXP kernel, SEXP state, SEXP stream) {
  fmcmc_model* m = model_of(model);
  fmcmc_run_spec rs;
  memset(&rs, 0, sizeof rs);
  rs.nsteps = (int64_t)asReal(elt(run, "nsteps"));
  rs.burnin = (int64_t)real_or(elt(run, "burnin"), 0);
  rs.thin = (int64_t)real_or(elt(run, "thin"), 1);
  rs.nchains = asInteger(elt(run, "nchains"));
  rs.flags = (uint32_t)int_or(elt(run, "flags"), 0) | FMCMC_RUN_COLMAJOR;
  rs.chain_offset = (int64_t)real_or(elt(run, "chain_offset"), 0);
  rs.nchains_total = (int64_t)real_or(elt(run, "nchains_total"), 0);
  rs.initial = dbl_or_null(elt(run, "initial")); /* k x nchains in R == [nchains][k] row-major */
  if (rs.nchains < 1) error("`nchains` must be an integer greater than 1.");

  fmcmc_kernel_spec ks;
  memset(&ks, 0, sizeof ks);
  ks.type = asInteger(elt(kernel, "type"));
  ks.k = asInteger(elt(kernel, "k"));
  ks.scheme = int_or(elt(kernel, "scheme"), FMCMC_SCHEME_JOINT);
  SEXP ord = elt(kernel, "order");
  ks.order = isNull(ord) ? NULL : INTEGER(ord);
  ks.order_len = isNull(ord) ? 0 : LENGTH(ord);
  ks.mu = dbl_or_null(elt(kernel, "mu"));
  ks.scale = dbl_or_null(elt(kernel, "scale"));
  ks.min_ = dbl_or_null(elt(kernel, "min"));
  ks.max_ = dbl_or_null(elt(kernel, "max"));
  ks.lb = dbl_or_null(elt(kernel, "lb"));
  ks.ub = dbl_or_null(elt(kernel, "ub"));
  SEXP fx = elt(kernel, "fixed");
  ks.fixed = isNull(fx) ? NULL : RAW(fx);
  ks.warmup = (int64_t)real_or(elt(kernel, "warmup"), 0);
  ks.freq = (int64_t)real_or(elt(kernel, "freq"), 1);
  ks.bw = (int64_t)real_or(elt(kernel, "bw"), 0);
  ks.until = real_or(elt(kernel, "until"), 1.0 / 0.0);
  ks.eps = real_or(elt(kernel, "eps"), 1e-4);
  ks.Sd = real_or(elt(kernel, "Sd"), -1.0);
  ks.arate = real_or(elt(kernel, "arate"), 0.234);
  ks.constr = dbl_or_null(elt(kernel, "constr"));
  ks.mvn_method = int_or(elt(kernel, "mvn_method"), FMCMC_MVN_CHOLESKY);
  /* nadapt arrives as double; the ABI wants int64 */
  SEXP na = elt(kernel, "nadapt");
  int64_t nadapt[64];
  ks.nadapt_len = isNull(na) ? 0 : LENGTH(na);
  if (ks.nadapt_len > 64) error("too many nadapt checkpoints (%d > 64)", ks.nadapt_len);
  for (int i = 0; i < ks.nadapt_len; i++) nadapt[i] = (int64_t)REAL(na)[i];
  ks.nadapt = ks.nadapt_len ? nadapt : NULL;
  if (ks.k < 1) error("the kernel has no parameters");

  int kf = 0;
  for (int j = 0; j < ks.k; j++) kf += !(ks.fixed && ks.fixed[j]);
  const int64_t dlen = fmcmc_kernel_state_len(ks.type, ks.k, kf);
  if (dlen < 0) error("unknown kernel type %d", ks.type);
  const int64_t keep = fmcmc_rows_kept(rs.nsteps, rs.burnin, rs.thin);
  if (keep < 0) error("-burnin- (%lld) cannot be >= than -nsteps- (%lld).", (long long)rs.burnin, (long long)rs.nsteps);

  /* state: copied, so the caller's objects are never modified in place (R semantics) */
  const R_xlen_t n_ist = (R_xlen_t)rs.nchains * FMCMC_ISTATE_LEN, n_dst = (R_xlen_t)rs.nchains * (dlen ? dlen : 1);
  SEXP ist_in = elt(state, "istate"), dst_in = elt(state, "dstate");
  if (!isNull(ist_in) && XLENGTH(ist_in) != n_ist) error("state$istate must have %lld elements", (long long)n_ist);
  if (dlen && !isNull(dst_in) && XLENGTH(dst_in) != n_dst) error("state$dstate must have %lld elements", (long long)n_dst);
  SEXP ist = PROTECT(allocVector(REALSXP, n_ist));
  SEXP dst = PROTECT(allocVector(REALSXP, n_dst));
  int64_t* ist64 = (int64_t*)R_alloc((size_t)n_ist, sizeof(int64_t)); /* transient: R reclaims it when the .Call returns */
  for (R_xlen_t i = 0; i < n_ist; i++) ist64[i] = isNull(ist_in) ? 0 : (int64_t)REAL(ist_in)[i];
  memset(REAL(dst), 0, (size_t)n_dst * sizeof(double));
  if (dlen && !isNull(dst_in)) memcpy(REAL(dst), REAL(dst_in), (size_t)n_dst * sizeof(double));
  fmcmc_kernel_state st = {ist64, dlen ? REAL(dst) : NULL};

  fmcmc_stream_spec sp;
  memset(&sp, 0, sizeof sp);
  sp.mode = int_or(elt(stream, "mode"), FMCMC_STREAM_PHILOX);
  sp.kdraw = int_or(elt(stream, "kdraw"), 0);
  sp.seed = (uint64_t)real_or(elt(stream, "seed"), 0);
  sp.run_index = (uint64_t)real_or(elt(stream, "run_index"), 0);
  sp.logu = dbl_or_null(elt(stream, "logu"));
  sp.z = dbl_or_null(elt(stream, "z"));

  const int want_draws = !(rs.flags & FMCMC_RUN_NO_DRAWS);
  SEXP ans = PROTECT(alloc3DArray(REALSXP, (int)keep, ks.k, rs.nchains));
  SEXP drw = PROTECT(want_draws ? alloc3DArray(REALSXP, (int)keep, ks.k, rs.nchains) : allocVector(REALSXP, 0));
  SEXP lp = PROTECT(allocMatrix(REALSXP, (int)keep, rs.nchains));
  fmcmc_run_report rep;
  memset(&rep, 0, sizeof rep);
  char err[1024] = "";
  const int rc = fmcmc_run(m, &rs, &ks, &st, &sp, REAL(ans), want_draws ? REAL(drw) : NULL, REAL(lp), &rep, err, sizeof err);
  for (R_xlen_t i = 0; i < n_ist; i++) REAL(ist)[i] = (double)ist64[i];
  if (rc) {
    UNPROTECT(5);
    error("%s", err); /* the library has already released its scratch */
  }
  SEXP out = PROTECT(allocVector(VECSXP, 6)), nm = PROTECT(allocVector(STRSXP, 6));
  static const char* names[] = {"ans", "draws", "logpost", "report", "istate", "dstate"};
  SEXP r = PROTECT(allocVector(REALSXP, 8));
  REAL(r)[0] = (double)rep.rows_kept;
  REAL(r)[1] = (double)rep.first_iter;
  REAL(r)[2] = (double)rep.last_iter;
  REAL(r)[3] = (double)rep.n_accept;
  REAL(r)[4] = rep.device_ms;
  REAL(r)[5] = (double)rep.path;
  REAL(r)[6] = (double)rep.n_launches;
  REAL(r)[7] = (double)rep.d2h_bytes;
  SEXP v[6];
  v[0] = ans; v[1] = drw; v[2] = lp; v[3] = r; v[4] = ist; v[5] = dst;
  for (int i = 0; i < 6; i++) {
    SET_VECTOR_ELT(out, i, v[i]);
    SET_STRING_ELT(nm, i, mkChar(names[i]));
  }
  setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(8);
  return out;
}

/* .Call(C_fmcmc_gelman, model, free_mask (raw), start_iter, thin) -> list(psrf, mpsrf, niter): coda's autoburnin window
 * and gelman.diag on everything the FMCMC_RUN_APPEND runs stored.  chol(W) failure (FMCMC_ENOTPD) -> mpsrf = NA, which the R
 * glue turns into the reference's warning + FALSE (R/convergence.R:207-217). */
SEXP C_fmcmc_gelman(SEXP model, SEXP free_mask, SEXP start_iter, SEXP thin) {
  fmcmc_model* m = model_of(model);
  int kf = 0;
  for (R_xlen_t j = 0; j < XLENGTH(free_mask); j++) kf += RAW(free_mask)[j] != 0;
  SEXP psrf = PROTECT(allocVector(REALSXP, kf));
  double mpsrf = NA_REAL;
  int64_t niter = 0;
  char err[1024] = "";
  const int rc = fmcmc_gelman(m, RAW(free_mask), (int64_t)asReal(start_iter), (int64_t)asReal(thin), REAL(psrf), &mpsrf, &niter,
                              err, sizeof err);
  if (rc && rc != FMCMC_ENOTPD) {
    UNPROTECT(1);
    error("%s", err);
  }
  SEXP out = PROTECT(allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, psrf);
  SET_VECTOR_ELT(out, 1, ScalarReal(rc == FMCMC_ENOTPD ? NA_REAL : mpsrf));
  SET_VECTOR_ELT(out, 2, ScalarReal((double)niter));
  UNPROTECT(2);
  return out;
}

/* .Call(C_fmcmc_cov_recursive, t(X_t), Mean_t_prev, Cov_t, t., eps, Sd, Ik or NULL, device) -> list(Mean_t, Cov_t)
 * (R/recursive.R:63-139; X_t arrives transposed: k x rows in R == [rows][k] row-major) */
SEXP C_fmcmc_cov_recursive(SEXP Xt, SEXP mean_prev, SEXP cov_prev, SEXP t, SEXP eps, SEXP Sd, SEXP Ik, SEXP device) {
  const int k = LENGTH(mean_prev);
  const int64_t rows = (int64_t)(XLENGTH(Xt) / (k > 0 ? k : 1));
  SEXP mo = PROTECT(allocVector(REALSXP, k)), co = PROTECT(allocMatrix(REALSXP, k, k));
  char err[1024] = "";
  const int rc = fmcmc_cov_recursive(asInteger(device), k, rows, REAL(Xt), REAL(mean_prev), REAL(cov_prev), asReal(t), asReal(eps),
                                     asReal(Sd), dbl_or_null(Ik), REAL(mo), REAL(co), err, sizeof err);
  if (rc) {
    UNPROTECT(2);
    error("%s", err);
  }
  SEXP out = PROTECT(allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, mo);
  SET_VECTOR_ELT(out, 1, co);
  UNPROTECT(3);
  return out;
}

/* .Call(C_fmcmc_reflect, x, lb, ub, which (raw mask or NULL), device) -> x reflected (R/kernel.R:450-493) */
SEXP C_fmcmc_reflect(SEXP x, SEXP lb, SEXP ub, SEXP which, SEXP device) {
  const int k = LENGTH(lb);
  SEXP out = PROTECT(allocVector(REALSXP, XLENGTH(x)));
  memcpy(REAL(out), REAL(x), (size_t)XLENGTH(x) * sizeof(double));
  char err[1024] = "";
  const int rc = fmcmc_reflect(asInteger(device), k, (int64_t)(XLENGTH(x) / (k > 0 ? k : 1)), REAL(out), REAL(lb), REAL(ub),
                               isNull(which) ? NULL : RAW(which), err, sizeof err);
  if (rc) {
    UNPROTECT(1);
    error("%s", err);
  }
  UNPROTECT(1);
  return out;
}

static const R_CallMethodDef calls[] = {
    {"C_fmcmc_model_create", (DL_FUNC)&C_fmcmc_model_create, 8},
    {"C_fmcmc_model_free", (DL_FUNC)&C_fmcmc_model_free, 1},
    {"C_fmcmc_store_reset", (DL_FUNC)&C_fmcmc_store_reset, 4},
    {"C_fmcmc_run", (DL_FUNC)&C_fmcmc_run, 5},
    {"C_fmcmc_gelman", (DL_FUNC)&C_fmcmc_gelman, 4},
    {"C_fmcmc_cov_recursive", (DL_FUNC)&C_fmcmc_cov_recursive, 8},
    {"C_fmcmc_reflect", (DL_FUNC)&C_fmcmc_reflect, 5},
    {NULL, NULL, 0}};
void R_init_fmcmc(DllInfo* dll) {
  R_registerRoutines(dll, NULL, calls, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
