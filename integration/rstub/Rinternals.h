/*
 * rstub/Rinternals.h - a MINIMAL stand-in for R's C API (test infrastructure, not R).
 *
 * R is not installed in the build image or on the GPU box, so integration/src/shim.c (the .Call layer a maintainer
 * adds to the reference package) could never see a compiler.  This stub declares exactly the subset of R's API the
 * shim uses, with R's names, argument orders and semantics (tagged-struct SEXP, names attribute, dim attribute,
 * external pointers with finalizers, PROTECT bookkeeping checked for balance, Rf_error as a longjmp like R's), so that
 * tests/test_shim_compile.py can compile the shim with -Wall -Werror against include/fmcmc_b200.h and
 * integration/test/drive_shim.c can call C_fmcmc_model_create / C_fmcmc_run / C_fmcmc_gelman with hand-built lists, the
 * way R/device.R does.  With real R the shim is compiled against R's own headers instead (src/Makevars).
 */
#ifndef RSTUB_RINTERNALS_H
#define RSTUB_RINTERNALS_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef ptrdiff_t R_xlen_t;
typedef unsigned char Rbyte;
typedef enum { FALSE = 0, TRUE = 1 } Rboolean;

enum { NILSXP = 0, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, EXTPTRSXP = 22, RAWSXP = 24, CHARSXP = 9 };

typedef struct SEXPREC* SEXP;
typedef void (*R_CFinalizer_t)(SEXP);
struct SEXPREC {
  int type;
  R_xlen_t length;
  void* data;          /* double / int / Rbyte / SEXP payload; char* for CHARSXP; the address for EXTPTRSXP */
  SEXP names;          /* names attribute (STRSXP) or R_NilValue */
  SEXP dim;            /* dim attribute (INTSXP) or R_NilValue   */
  R_CFinalizer_t fin;  /* EXTPTRSXP only */
};

extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;
extern double R_NaReal;
#define NA_REAL R_NaReal

SEXP Rf_allocVector(int type, R_xlen_t n);
SEXP Rf_allocMatrix(int type, int nrow, int ncol);
SEXP Rf_alloc3DArray(int type, int nrow, int ncol, int nface);
SEXP Rf_getAttrib(SEXP x, SEXP what);
SEXP Rf_setAttrib(SEXP x, SEXP what, SEXP value);
SEXP Rf_mkChar(const char* s);
SEXP Rf_ScalarReal(double v);
SEXP Rf_ScalarInteger(int v);
int Rf_asInteger(SEXP x);
double Rf_asReal(SEXP x);
int Rf_ncols(SEXP x);
int Rf_nrows(SEXP x);
int Rf_isNull(SEXP x);
R_xlen_t Rf_xlength(SEXP x);
int Rf_length(SEXP x);
#ifdef __GNUC__
void Rf_error(const char* fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
#else
void Rf_error(const char* fmt, ...);
#endif

double* REAL(SEXP x);
int* INTEGER(SEXP x);
Rbyte* RAW(SEXP x);
SEXP STRING_ELT(SEXP x, R_xlen_t i);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v);
const char* R_CHAR(SEXP x);
#define CHAR(x) R_CHAR(x)

SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot);
void* R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);

char* R_alloc(size_t n, int size); /* transient storage, reclaimed when the .Call returns (here: by rstub_end_call) */
SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)

/* the short names R.h / Rinternals.h define unless R_NO_REMAP is set */
#define allocVector Rf_allocVector
#define allocMatrix Rf_allocMatrix
#define alloc3DArray Rf_alloc3DArray
#define getAttrib Rf_getAttrib
#define setAttrib Rf_setAttrib
#define mkChar Rf_mkChar
#define ScalarReal Rf_ScalarReal
#define ScalarInteger Rf_ScalarInteger
#define asInteger Rf_asInteger
#define asReal Rf_asReal
#define ncols Rf_ncols
#define nrows Rf_nrows
#define isNull Rf_isNull
#define XLENGTH(x) Rf_xlength(x)
#define LENGTH(x) Rf_length(x)
#define error Rf_error

/* ---- test-side helpers (NOT part of R's API): what the R interpreter itself does around a .Call ---- */
#include <setjmp.h>
extern jmp_buf rstub_error_jmp;       /* Rf_error longjmps here once rstub_error_armed is set (else it aborts) */
extern int rstub_error_armed;
extern char rstub_error_msg[2048];
void rstub_end_call(void);            /* frees R_alloc'ed storage, like R does on return from .Call */
int rstub_protect_depth(void);        /* must be back to 0 after every .Call (R warns "stack imbalance") */
void rstub_run_finalizers(void);      /* gc at exit: runs every registered external-pointer finalizer once */
SEXP rstub_named_list(int n, const char** names, SEXP* values);
SEXP rstub_real(R_xlen_t n, const double* v);
SEXP rstub_raw(R_xlen_t n, const unsigned char* v);
SEXP rstub_int(R_xlen_t n, const int* v);

#ifdef __cplusplus
}
#endif
#endif
