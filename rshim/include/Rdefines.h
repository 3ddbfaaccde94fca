/* rshim: subset of R's <Rdefines.h> (see Rinternals.h in this directory). */
#ifndef RSHIM_RDEFINES_H
#define RSHIM_RDEFINES_H

#include "Rinternals.h"

#define NEW_LOGICAL(n)   Rf_allocVector(LGLSXP, n)
#define NEW_INTEGER(n)   Rf_allocVector(INTSXP, n)
#define NEW_NUMERIC(n)   Rf_allocVector(REALSXP, n)
#define NEW_COMPLEX(n)   Rf_allocVector(CPLXSXP, n)
#define NEW_CHARACTER(n) Rf_allocVector(STRSXP, n)
#define NEW_RAW(n)       Rf_allocVector(RAWSXP, n)
#define NEW_LIST(n)      Rf_allocVector(VECSXP, n)

#define IS_LOGICAL(x)    (TYPEOF(x) == LGLSXP)
#define IS_INTEGER(x)    (TYPEOF(x) == INTSXP)
#define IS_NUMERIC(x)    (TYPEOF(x) == REALSXP)
#define IS_COMPLEX(x)    (TYPEOF(x) == CPLXSXP)
#define IS_CHARACTER(x)  (TYPEOF(x) == STRSXP)
#define IS_RAW(x)        (TYPEOF(x) == RAWSXP)
#define IS_LIST(x)       (TYPEOF(x) == VECSXP)

#define GET_DIM(x)          Rf_getDim(x)
#define SET_DIM(x, v)       Rf_setDim(x, v)
#define GET_NAMES(x)        Rf_getNames(x)
#define SET_NAMES(x, v)     Rf_setNames(x, v)
#define GET_DIMNAMES(x)     Rf_getDimnames(x)
#define SET_DIMNAMES(x, v)  Rf_setDimnames(x, v)
#define GET_LENGTH(x)       LENGTH(x)

#endif  /* RSHIM_RDEFINES_H */
