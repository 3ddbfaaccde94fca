/* TEST INFRASTRUCTURE -- glue needed to link the reference's hot-path C files
 * (compiled in place from /root/reference/src, see Makefile) into
 * oracle/_ref/libsvtref.so without the rest of the package.
 *
 *  - the R / S4Vectors API entries declared in ref_extra.h (see there).
 *  - svtref_init(): what R_init_SparseArray() does to the NA globals,
 *    src/R_init_SparseArray.c:153-154 (the registration table itself names
 *    every .Call entry point of the package, so that file cannot be linked).
 */
#include <Rdefines.h>
#include <stdlib.h>
#include <string.h>
#include "ref_extra.h"

extern int intNA;
extern double doubleNA;
extern Rcomplex RcomplexNA;

/* any attribute at all?  (the shim models dim / names / dimnames) */
SEXP svtref_attrib(SEXP x)
{
	if (x->dim != R_NilValue) return x->dim;
	if (x->names != R_NilValue) return x->names;
	return x->dimnames;
}

SEXP svtref_get_class(SEXP x)
{
	Rf_error("S4 classes are not modelled by the R shim");
	return R_NilValue;
}

double R_strtod(const char *c, char **end) { return strtod(c, end); }

Rboolean StringTrue(const char *name)
{
	static const char *t[] = {"T", "True", "TRUE", "true"};
	for (int i = 0; i < 4; i++)
		if (strcmp(name, t[i]) == 0) return TRUE;
	return FALSE;
}

Rboolean StringFalse(const char *name)
{
	static const char *f[] = {"F", "False", "FALSE", "false"};
	for (int i = 0; i < 4; i++)
		if (strcmp(name, f[i]) == 0) return TRUE;
	return FALSE;
}

SEXP Rf_coerceVector(SEXP v, SEXPTYPE type)
{
	Rf_error("coerceVector() to/from lists and strings is not modelled "
		 "by the R shim");
	return R_NilValue;
}

/* stable merge sort of the order vector by key */
int sort_ints(int *base, int base_len, const int *x, int desc, int use_radix,
	      unsigned short int *rxbuf1, int *rxbuf2)
{
	int sorted = 1;
	for (int i = 1; i < base_len && sorted; i++)
		if (desc ? x[base[i - 1]] < x[base[i]]
			 : x[base[i - 1]] > x[base[i]])
			sorted = 0;
	if (sorted)
		return 0;
	int *tmp = (int *) malloc(sizeof(int) * (size_t) base_len);
	if (tmp == NULL)
		return -1;
	for (int w = 1; w < base_len; w *= 2) {
		for (int lo = 0; lo < base_len; lo += 2 * w) {
			int mid = lo + w < base_len ? lo + w : base_len;
			int hi = lo + 2 * w < base_len ? lo + 2 * w : base_len;
			int a = lo, b = mid, k = lo;
			while (a < mid && b < hi) {
				int take_b = desc ? x[base[b]] > x[base[a]]
						  : x[base[b]] < x[base[a]];
				tmp[k++] = take_b ? base[b++] : base[a++];
			}
			while (a < mid) tmp[k++] = base[a++];
			while (b < hi) tmp[k++] = base[b++];
		}
		memcpy(base, tmp, sizeof(int) * (size_t) base_len);
	}
	free(tmp);
	return 1;
}

__attribute__((constructor)) static void svtref_init(void)
{
	intNA = NA_INTEGER;
	doubleNA = RcomplexNA.r = RcomplexNA.i = NA_REAL;
}
