/* rstub/R.h - see Rinternals.h in this directory: a minimal stand-in for R's C API (test infrastructure only). */
#ifndef RSTUB_R_H
#define RSTUB_R_H
#include <stdlib.h>
#include <string.h>
#include "Rinternals.h"
#endif
