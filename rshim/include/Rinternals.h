/* rshim -- a minimal stand-in for R's C API (libR + <Rinternals.h>).
 *
 * R is not installed in the build container nor on the GPU boxes, so neither
 * the reference package (Bioconductor/SparseArray, plain C against R's API)
 * nor our own R-facing glue (sparsearray_b200/rglue/) can be compiled against
 * the real headers here.  This shim implements the small subset of the API
 * that the SVT hot path touches, with R's exact NA encodings
 * (NA_INTEGER = INT_MIN, NA_REAL = NaN with low word 1954), so that
 *   (a) the reference's own C sources compile unmodified into oracle/_ref/,
 *   (b) our glue compiles unmodified here and against real R elsewhere.
 * It is NOT a product component: with real R the glue links against libR.
 *
 * Object model: every SEXP is a heap record {type, length, attribs, data}.
 * No garbage collector: objects live until rshim_release()/rshim_release_tree().
 * Vectors can wrap caller-owned memory (rshim_wrap_vector) for zero-copy use
 * of numpy buffers.
 */
#ifndef RSHIM_RINTERNALS_H
#define RSHIM_RINTERNALS_H

#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned int SEXPTYPE;
typedef ptrdiff_t R_xlen_t;
typedef unsigned char Rbyte;
typedef struct { double r; double i; } Rcomplex;
typedef enum { FALSE = 0, TRUE } Rboolean;

#define NILSXP	     0
#define SYMSXP	     1
#define LISTSXP	     2
#define CHARSXP	     9
#define LGLSXP	    10
#define INTSXP	    13
#define REALSXP	    14
#define CPLXSXP	    15
#define STRSXP	    16
#define VECSXP	    19
#define EXTPTRSXP   22
#define RAWSXP	    24

typedef struct SEXPREC {
	SEXPTYPE type;
	int owns_data;          /* 1: data malloc'ed by the shim, 0: wrapped */
	R_xlen_t length;
	void *data;             /* int*, double*, Rcomplex*, Rbyte*, SEXP*, char* */
	struct SEXPREC *dim;    /* "dim" attribute or R_NilValue */
	struct SEXPREC *names;  /* "names" attribute or R_NilValue */
	struct SEXPREC *dimnames;
	void (*finalizer)(struct SEXPREC *);   /* EXTPTRSXP only, may be NULL */
} SEXPREC, *SEXP;

extern SEXP R_NilValue;
extern SEXP R_NaString;
extern SEXP R_BlankString;
extern int R_NaInt;
extern double R_NaReal;
extern double R_NaN;
extern double R_PosInf;
extern double R_NegInf;

#define NA_INTEGER R_NaInt
#define NA_LOGICAL R_NaInt
#define NA_REAL    R_NaReal
#define NA_STRING  R_NaString

int R_IsNA(double x);
int R_IsNaN(double x);
int R_finite(double x);
#define ISNAN(x)    (isnan(x) != 0)
#define ISNA(x)     R_IsNA(x)
#define R_FINITE(x) R_finite(x)

/* accessors */
#define TYPEOF(x)   ((x)->type)
#define XLENGTH(x)  ((x)->length)
#define LENGTH(x)   ((int) (x)->length)
#define length(x)   LENGTH(x)
#define xlength(x)  XLENGTH(x)
#define INTEGER(x)  ((int *) (x)->data)
#define LOGICAL(x)  ((int *) (x)->data)
#define REAL(x)     ((double *) (x)->data)
#define COMPLEX(x)  ((Rcomplex *) (x)->data)
#define RAW(x)      ((Rbyte *) (x)->data)
#define DATAPTR(x)  ((x)->data)
#define CHAR(x)     ((const char *) (x)->data)
#define VECTOR_ELT(x, i)  (((SEXP *) (x)->data)[i])
#define STRING_ELT(x, i)  (((SEXP *) (x)->data)[i])
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v);

#define PROTECT(x)   (x)
#define UNPROTECT(n) ((void) (n))

/* allocation */
SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n);
SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol);
SEXP Rf_allocArray(SEXPTYPE type, SEXP dims);
SEXP Rf_duplicate(SEXP x);
SEXP Rf_mkChar(const char *s);
SEXP Rf_mkString(const char *s);
SEXP Rf_ScalarInteger(int v);
SEXP Rf_ScalarLogical(int v);
SEXP Rf_ScalarReal(double v);
SEXP Rf_ScalarString(SEXP v);
char *R_alloc(size_t n, int size);

#define allocVector   Rf_allocVector
#define allocMatrix   Rf_allocMatrix
#define allocArray    Rf_allocArray
#define duplicate     Rf_duplicate
#define mkChar        Rf_mkChar
#define mkString      Rf_mkString
#define ScalarInteger Rf_ScalarInteger
#define ScalarLogical Rf_ScalarLogical
#define ScalarReal    Rf_ScalarReal
#define ScalarString  Rf_ScalarString

/* external pointers: `data` is the address; there is no garbage collector in
   the shim, so a registered C finalizer runs when the record is released
   (rshim_release), which is where R would run it at the latest */
typedef void (*R_CFinalizer_t)(SEXP);
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);

/* predicates / misc */
int Rf_isVectorList(SEXP x);
int Rf_isBlankString(const char *s);
const char *Rf_type2char(SEXPTYPE t);
SEXPTYPE Rf_str2type(const char *s);
#define isVectorList  Rf_isVectorList
#define isBlankString Rf_isBlankString
#define type2char     Rf_type2char
#define str2type      Rf_str2type
#define isNull(x)     ((x) == R_NilValue)

/* attributes (only dim / names / dimnames are modelled) */
SEXP Rf_getDim(SEXP x);
SEXP Rf_setDim(SEXP x, SEXP v);
SEXP Rf_getNames(SEXP x);
SEXP Rf_setNames(SEXP x, SEXP v);
SEXP Rf_getDimnames(SEXP x);
SEXP Rf_setDimnames(SEXP x, SEXP v);

/* conditions */
void Rf_error(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
void Rf_warning(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
#define error   Rf_error
#define warning Rf_warning

/* ---- shim-only API (used by the Python/C test harnesses) ---- */

/* Wrap caller-owned memory as an R vector (zero-copy). */
SEXP rshim_wrap_vector(SEXPTYPE type, R_xlen_t n, void *data);
void rshim_release(SEXP x);       /* free one record (+ owned payload) */
void rshim_release_tree(SEXP x);  /* recursive over VECSXP/STRSXP/attribs */

/* Build list(nzvals, nzoffs) leaves for columns [0, ncol) of a CSC matrix,
   wrapping the CSC arrays in place.  'lacunar' (may be NULL) flags leaves
   whose nzvals must be NULL.  Empty columns become R_NilValue.  Returns
   R_NilValue when nnz == 0 (as SVT_SparseArray objects do). */
SEXP rshim_svt_from_csc(int ncol, const int64_t *ptr, int *offs, void *vals,
			SEXPTYPE vals_type, const unsigned char *lacunar);

/* Call 'fn' (a .Call entry point taking 'nargs' SEXPs) catching error().
   Returns NULL and sets *status=1 on error; message via rshim_last_error().
   Memory obtained with R_alloc() during the call is released on return. */
SEXP rshim_try_call(void *fn, int nargs, SEXP *args, int *status);
const char *rshim_last_error(void);
int rshim_warning_count(void);
const char *rshim_warning_message(int i);
void rshim_clear_warnings(void);
long rshim_live_objects(void);

#ifdef __cplusplus
}
#endif

#endif  /* RSHIM_RINTERNALS_H */
