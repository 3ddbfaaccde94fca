/* TEST INFRASTRUCTURE -- force-included (gcc -include) when the reference's
 * sources are compiled into oracle/_ref/libsvtref.so.  Declares the few R /
 * S4Vectors API entries that src/leaf_utils.c, src/coerceVector2.c and
 * src/SVT_SparseArray_class.c reference but the R shim does not model; they
 * are defined in ref_support.c.  None of them is reached by the calls the
 * oracle makes (transpose, SVT <-> CSC, statistics, products) except
 * sort_ints(), restated from the documented behaviour of S4Vectors'
 * sort_ints(): order `base` by x[base[i]] (ascending, stable); return 0 when
 * the order was already sorted, 1 when it changed, < 0 on error.
 */
#ifndef SVT_REF_EXTRA_H
#define SVT_REF_EXTRA_H
#include <Rdefines.h>

SEXP svtref_attrib(SEXP x);
#define ATTRIB(x) svtref_attrib(x)
SEXP svtref_get_class(SEXP x);
#define GET_CLASS(x) svtref_get_class(x)
double R_strtod(const char *c, char **end);
Rboolean StringTrue(const char *name);
Rboolean StringFalse(const char *name);
SEXP Rf_coerceVector(SEXP v, SEXPTYPE type);
#define coerceVector Rf_coerceVector
int sort_ints(int *base, int base_len, const int *x, int desc, int use_radix,
	      unsigned short int *rxbuf1, int *rxbuf2);

#endif
