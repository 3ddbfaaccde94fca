/* TEST INFRASTRUCTURE -- glue needed to link the reference's hot-path C files
 * (compiled in place from /root/reference/src, see Makefile) into
 * oracle/_ref/libsvtref.so without the rest of the package.
 *
 *  - _REC_nzcount_SVT(): restated from src/SVT_SparseArray_class.c:200-218
 *    (that file cannot be compiled: it needs the un-vendored S4Vectors
 *    headers); the only caller on the path is src/SparseMatrix_mult.c.
 *  - svtref_init(): what R_init_SparseArray() does to the NA globals,
 *    src/R_init_SparseArray.c:153-154 (the registration table itself names
 *    every .Call entry point of the package, so that file cannot be linked).
 */
#include <Rdefines.h>

extern int intNA;
extern double doubleNA;
extern Rcomplex RcomplexNA;

R_xlen_t _REC_nzcount_SVT(SEXP SVT, int ndim)
{
	if (SVT == R_NilValue)
		return 0;
	if (ndim == 1) {
		SEXP nzoffs = VECTOR_ELT(SVT, 1);
		return XLENGTH(nzoffs);
	}
	R_xlen_t nzcount = 0;
	int n = LENGTH(SVT);
	for (int i = 0; i < n; i++)
		nzcount += _REC_nzcount_SVT(VECTOR_ELT(SVT, i), ndim - 1);
	return nzcount;
}

__attribute__((constructor)) static void svtref_init(void)
{
	intNA = NA_INTEGER;
	doubleNA = RcomplexNA.r = RcomplexNA.i = NA_REAL;
}
