/* rstub/R_ext/Rdynload.h - registration types of R's dynamic loader (see ../Rinternals.h: test infrastructure only). */
#ifndef RSTUB_RDYNLOAD_H
#define RSTUB_RDYNLOAD_H
#include "../Rinternals.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct { const char* name; DL_FUNC fun; int numArgs; void* types; } R_CMethodDef;
typedef R_CallMethodDef R_ExternalMethodDef;
typedef R_CMethodDef R_FortranMethodDef;
typedef struct DllInfo { const R_CallMethodDef* calls; int dynamic_symbols; } DllInfo;
int R_registerRoutines(DllInfo* info, const R_CMethodDef* c, const R_CallMethodDef* call, const R_FortranMethodDef* f,
                       const R_ExternalMethodDef* ext);
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value);
#ifdef __cplusplus
}
#endif
#endif
