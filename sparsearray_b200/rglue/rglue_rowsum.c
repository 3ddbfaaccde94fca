/* .Call entry points C_rowsum_SVT / C_colsum_SVT served by the GPU path:
 * rowsum() / colsum() of an SVT_SparseMatrix (R/rowsum-methods.R:7-49).
 *
 * Same signatures, argument checks, error messages, result type (integer for
 * integer input, double for double input), zero-filled result for a NULL SVT
 * and "NAs produced by integer overflow" warning as the reference
 * (src/rowsum_methods.c:15-37, :281-326, :364-409); the per-leaf loops are
 * replaced by svtgpu_rowsum() / svtgpu_colsum().  The dgCMatrix variants are
 * not served (they take Matrix objects, not SVTs).
 */
#include "rglue_common.h"

#include <limits.h>
#include <string.h>

/* check_group(), src/rowsum_methods.c:15-37 */
static void check_group(SEXP group, int x_nrow, int ngroup)
{
	if (!IS_INTEGER(group))
		error("the grouping vector must be "
		      "an integer vector or factor");
	if (LENGTH(group) != x_nrow)
		error("the grouping vector must have one element "
		      "per row in 'x' for rowsum()\n  and one element "
		      "per column in 'x' for colsum()");
	for (int i = 0; i < x_nrow; i++) {
		int g = INTEGER(group)[i];
		if (g == NA_INTEGER) {
			if (ngroup < 1)
				error("'ngroup' must be >= 1 when 'group' "
				      "contains missing values");
		} else {
			if (g < 1 || g > ngroup)
				error("all non-NA values in 'group' must "
				      "be >= 1 and <= 'ngroup'");
		}
	}
}

static SEXP groupsum(SEXP x_dim, SEXP x_type, SEXP x_SVT, SEXP group,
		     SEXP ngroup, SEXP na_rm, int by_row, const char *fun)
{
	if (LENGTH(x_dim) != 2)
		error("input object must have 2 dimensions");
	int x_nrow = INTEGER(x_dim)[0];
	int x_ncol = INTEGER(x_dim)[1];
	int narm = LOGICAL(na_rm)[0];
	SEXPTYPE x_Rtype = rglue_get_and_check_Rtype(x_type, fun, "x_type");
	int ng = INTEGER(ngroup)[0];
	check_group(group, by_row ? x_nrow : x_ncol, ng);
	int ans_nrow = by_row ? ng : x_nrow;
	int ans_ncol = by_row ? x_ncol : ng;
	if ((double) ans_nrow * (double) ans_ncol > INT_MAX)
		error("too many groups (matrix of sums will be too big)");
	if (x_Rtype != REALSXP && x_Rtype != INTSXP)
		error("rowsum() and colsum() do not support "
		      "SVT_SparseMatrix objects of\n"
		      "  type \"%s\" at the moment", type2char(x_Rtype));
	SEXP ans = PROTECT(allocMatrix(x_Rtype, ans_nrow, ans_ncol));
	memset(DATAPTR(ans), 0, (x_Rtype == REALSXP ? sizeof(double)
						    : sizeof(int)) *
				(size_t) XLENGTH(ans));
	if (x_SVT != R_NilValue && XLENGTH(ans) != 0 && x_nrow != 0 &&
	    x_ncol != 0) {
		rglue_input in;
		rglue_acquire(x_SVT, INTEGER(x_dim), 2, x_Rtype, 1, 1, &in);
		int overflow = 0;
		int rc = by_row
			? svtgpu_rowsum(in.m, INTEGER(group), ng, narm,
					DATAPTR(ans), &overflow)
			: svtgpu_colsum(in.m, INTEGER(group), ng, narm,
					DATAPTR(ans), &overflow);
		rglue_done(&in, fun);
		if (rc != SVTGPU_OK)
			rglue_fail(rc, by_row ? "svtgpu_rowsum"
					      : "svtgpu_colsum");
		if (overflow)
			warning("NAs produced by integer overflow");
	}
	UNPROTECT(1);
	return ans;
}

/* --- .Call ENTRY POINT --- */
SEXP C_rowsum_SVT(SEXP x_dim, SEXP x_type, SEXP x_SVT,
		  SEXP group, SEXP ngroup, SEXP na_rm)
{
	return groupsum(x_dim, x_type, x_SVT, group, ngroup, na_rm, 1,
			"C_rowsum_SVT");
}

/* --- .Call ENTRY POINT --- */
SEXP C_colsum_SVT(SEXP x_dim, SEXP x_type, SEXP x_SVT,
		  SEXP group, SEXP ngroup, SEXP na_rm)
{
	return groupsum(x_dim, x_type, x_SVT, group, ngroup, na_rm, 0,
			"C_colsum_SVT");
}
