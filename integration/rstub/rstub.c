#define _POSIX_C_SOURCE 200809L
/* rstub.c - implementation of the minimal R C-API stand-in declared in Rinternals.h (test infrastructure only). */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "R_ext/Rdynload.h"
#include "Rinternals.h"

static struct SEXPREC nil_rec = {NILSXP, 0, NULL, NULL, NULL, NULL};
static struct SEXPREC names_sym = {CHARSXP, 5, (void*)"names", NULL, NULL, NULL};
static struct SEXPREC dim_sym = {CHARSXP, 3, (void*)"dim", NULL, NULL, NULL};
SEXP R_NilValue = &nil_rec, R_NamesSymbol = &names_sym, R_DimSymbol = &dim_sym;
double R_NaReal = NAN;

jmp_buf rstub_error_jmp;
int rstub_error_armed = 0;
char rstub_error_msg[2048];
static int protect_depth = 0;
static SEXP extptrs[256];
static int n_extptrs = 0;

static size_t elt_size(int type) {
  switch (type) {
    case REALSXP: return sizeof(double);
    case INTSXP: case LGLSXP: return sizeof(int);
    case RAWSXP: return 1;
    case STRSXP: case VECSXP: return sizeof(SEXP);
    default: return 0;
  }
}
SEXP Rf_allocVector(int type, R_xlen_t n) {
  SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
  s->type = type;
  s->length = n;
  s->names = s->dim = R_NilValue;
  const size_t es = elt_size(type);
  if (es && n > 0) s->data = calloc((size_t)n, es);
  if (type == STRSXP || type == VECSXP)
    for (R_xlen_t i = 0; i < n; i++) ((SEXP*)s->data)[i] = R_NilValue;
  return s;
}
static SEXP with_dim(SEXP s, int nd, const int* d) {
  s->dim = Rf_allocVector(INTSXP, nd);
  memcpy(s->dim->data, d, (size_t)nd * sizeof(int));
  return s;
}
SEXP Rf_allocMatrix(int type, int nrow, int ncol) {
  const int d[2] = {nrow, ncol};
  return with_dim(Rf_allocVector(type, (R_xlen_t)nrow * ncol), 2, d);
}
SEXP Rf_alloc3DArray(int type, int nrow, int ncol, int nface) {
  const int d[3] = {nrow, ncol, nface};
  return with_dim(Rf_allocVector(type, (R_xlen_t)nrow * ncol * nface), 3, d);
}
SEXP Rf_getAttrib(SEXP x, SEXP what) {
  if (x == R_NilValue) return R_NilValue;
  if (what == R_NamesSymbol) return x->names;
  if (what == R_DimSymbol) return x->dim;
  return R_NilValue;
}
SEXP Rf_setAttrib(SEXP x, SEXP what, SEXP value) {
  if (what == R_NamesSymbol) x->names = value;
  else if (what == R_DimSymbol) x->dim = value;
  return value;
}
SEXP Rf_mkChar(const char* s) {
  SEXP c = (SEXP)calloc(1, sizeof(struct SEXPREC));
  c->type = CHARSXP;
  c->length = (R_xlen_t)strlen(s);
  c->data = strdup(s);
  c->names = c->dim = R_NilValue;
  return c;
}
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); ((double*)s->data)[0] = v; return s; }
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); ((int*)s->data)[0] = v; return s; }
int Rf_isNull(SEXP x) { return x == R_NilValue || x->type == NILSXP; }
R_xlen_t Rf_xlength(SEXP x) { return Rf_isNull(x) ? 0 : x->length; }
int Rf_length(SEXP x) { return (int)Rf_xlength(x); }
static void need(SEXP x, int type, const char* what) {
  if (x == NULL || x->type != type) Rf_error("%s() applied to an object of type %d", what, x ? x->type : -1);
}
double* REAL(SEXP x) { need(x, REALSXP, "REAL"); return (double*)x->data; }
int* INTEGER(SEXP x) { if (x && x->type == LGLSXP) return (int*)x->data; need(x, INTSXP, "INTEGER"); return (int*)x->data; }
Rbyte* RAW(SEXP x) { need(x, RAWSXP, "RAW"); return (Rbyte*)x->data; }
SEXP STRING_ELT(SEXP x, R_xlen_t i) { need(x, STRSXP, "STRING_ELT"); return ((SEXP*)x->data)[i]; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { need(x, VECSXP, "VECTOR_ELT"); return ((SEXP*)x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { need(x, VECSXP, "SET_VECTOR_ELT"); ((SEXP*)x->data)[i] = v; return v; }
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) { need(x, STRSXP, "SET_STRING_ELT"); ((SEXP*)x->data)[i] = v; }
const char* R_CHAR(SEXP x) { need(x, CHARSXP, "CHAR"); return (const char*)x->data; }
/* asInteger / asReal coerce the first element like R does (NA for NULL / empty) */
int Rf_asInteger(SEXP x) {
  if (Rf_xlength(x) < 1) return INT32_MIN;
  switch (x->type) {
    case INTSXP: case LGLSXP: return ((int*)x->data)[0];
    case REALSXP: { const double v = ((double*)x->data)[0]; return isnan(v) ? INT32_MIN : (int)v; }
    case RAWSXP: return ((Rbyte*)x->data)[0];
    default: return INT32_MIN;
  }
}
double Rf_asReal(SEXP x) {
  if (Rf_xlength(x) < 1) return R_NaReal;
  switch (x->type) {
    case INTSXP: case LGLSXP: return ((int*)x->data)[0] == INT32_MIN ? R_NaReal : (double)((int*)x->data)[0];
    case REALSXP: return ((double*)x->data)[0];
    case RAWSXP: return ((Rbyte*)x->data)[0];
    default: return R_NaReal;
  }
}
int Rf_nrows(SEXP x) { return (!Rf_isNull(x) && x->dim != R_NilValue) ? ((int*)x->dim->data)[0] : (int)Rf_xlength(x); }
int Rf_ncols(SEXP x) { return (!Rf_isNull(x) && x->dim != R_NilValue && x->dim->length >= 2) ? ((int*)x->dim->data)[1] : 1; }

void Rf_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(rstub_error_msg, sizeof rstub_error_msg, fmt, ap);
  va_end(ap);
  protect_depth = 0; /* R unwinds the protect stack to the context it jumps to */
  if (rstub_error_armed) longjmp(rstub_error_jmp, 1);
  fprintf(stderr, "Error: %s\n", rstub_error_msg);
  abort();
}

SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) {
  (void)tag; (void)prot;
  SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
  s->type = EXTPTRSXP;
  s->data = p;
  s->names = s->dim = R_NilValue;
  return s;
}
void* R_ExternalPtrAddr(SEXP s) { need(s, EXTPTRSXP, "R_ExternalPtrAddr"); return s->data; }
void R_ClearExternalPtr(SEXP s) { need(s, EXTPTRSXP, "R_ClearExternalPtr"); s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit) {
  (void)onexit;
  need(s, EXTPTRSXP, "R_RegisterCFinalizerEx");
  s->fin = fun;
  if (n_extptrs < 256) extptrs[n_extptrs++] = s;
}
void rstub_run_finalizers(void) {
  for (int i = 0; i < n_extptrs; i++)
    if (extptrs[i]->fin) { R_CFinalizer_t f = extptrs[i]->fin; extptrs[i]->fin = NULL; f(extptrs[i]); }
  n_extptrs = 0;
}

static void* transient[64];
static int n_transient = 0;
char* R_alloc(size_t n, int size) {
  void* p = calloc(n ? n : 1, (size_t)size);
  if (n_transient < 64) transient[n_transient++] = p;
  return (char*)p;
}
void rstub_end_call(void) {
  for (int i = 0; i < n_transient; i++) free(transient[i]);
  n_transient = 0;
}
SEXP Rf_protect(SEXP x) { protect_depth++; return x; }
void Rf_unprotect(int n) {
  protect_depth -= n;
  if (protect_depth < 0) { fprintf(stderr, "rstub: UNPROTECT(%d): stack imbalance\n", n); abort(); }
}
int rstub_protect_depth(void) { return protect_depth; }

int R_registerRoutines(DllInfo* info, const R_CMethodDef* c, const R_CallMethodDef* call, const R_FortranMethodDef* f,
                       const R_ExternalMethodDef* ext) {
  (void)c; (void)f; (void)ext;
  info->calls = call;
  return 1;
}
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value) { info->dynamic_symbols = value; return TRUE; }

SEXP rstub_real(R_xlen_t n, const double* v) { SEXP s = Rf_allocVector(REALSXP, n); if (n) memcpy(s->data, v, (size_t)n * 8); return s; }
SEXP rstub_int(R_xlen_t n, const int* v) { SEXP s = Rf_allocVector(INTSXP, n); if (n) memcpy(s->data, v, (size_t)n * 4); return s; }
SEXP rstub_raw(R_xlen_t n, const unsigned char* v) { SEXP s = Rf_allocVector(RAWSXP, n); if (n) memcpy(s->data, v, (size_t)n); return s; }
SEXP rstub_named_list(int n, const char** names, SEXP* values) {
  SEXP l = Rf_allocVector(VECSXP, n), nm = Rf_allocVector(STRSXP, n);
  for (int i = 0; i < n; i++) { SET_VECTOR_ELT(l, i, values[i]); SET_STRING_ELT(nm, i, Rf_mkChar(names[i])); }
  Rf_setAttrib(l, R_NamesSymbol, nm);
  return l;
}
