/* .Call entry points C_crossprod2_SVT_mat / C_crossprod2_mat_SVT served by
 * the GPU path: same signatures, checks and result as the reference's
 * (src/SparseMatrix_mult.c:931-1034); the K passes over the SVT
 * (crossprod2_SVT_mat_{double,int}(), :385-431,483-514) become one upload and
 * one gather kernel.  C_matmul_SVT_mat is an extension serving `svt %*% m`
 * without the reference's t(svt) (R/SparseMatrix-mult.R:196-198). */
#include "rglue_common.h"

#include <string.h>

static SEXPTYPE get_and_check_input_Rtype(SEXP type, const char *argname)
{
	SEXPTYPE Rtype = rglue_get_and_check_Rtype(type,
				"get_and_check_input_Rtype", argname);
	if (Rtype != REALSXP && Rtype != INTSXP)
		error("SparseArray internal error in "
		      "get_and_check_input_Rtype():\n"
		      "    input type \"%s\" is not supported yet",
		      type2char(Rtype));
	return Rtype;
}

static void check_ans_type(SEXP ans_type, const char *fun)
{
	SEXPTYPE ans_Rtype = rglue_get_and_check_Rtype(ans_type, fun,
						       "ans_type");
	if (ans_Rtype != REALSXP)
		error("SparseArray internal error in %s():\n"
		      "    output type \"%s\" is not supported yet",
		      fun, type2char(ans_Rtype));
}

/* _new_Rmatrix0(), src/Rvector_utils.c:485-496 */
static SEXP new_double_matrix0(int nrow, int ncol, SEXP dimnames)
{
	SEXP ans = PROTECT(allocMatrix(REALSXP, nrow, ncol));
	memset(REAL(ans), 0, sizeof(double) * (size_t) XLENGTH(ans));
	SET_DIMNAMES(ans, dimnames);
	UNPROTECT(1);
	return ans;
}

/* flatten + upload + product; 'svt_on_left' picks the output orientation */
static void run_crossprod(const int *svt_dim, SEXPTYPE Rtype, SEXP SVT,
			  SEXP dense, int dense_nrow, int dense_ncol,
			  int transpose_dense, int svt_on_left, double *out)
{
	rglue_input in;
	rglue_acquire(SVT, svt_dim, 2, Rtype, 1, 1, &in);
	int rc = svtgpu_crossprod(in.m, DATAPTR(dense), (int) Rtype,
				  dense_nrow, dense_ncol, transpose_dense,
				  svt_on_left, out);
	rglue_done(&in, svt_on_left ? "C_crossprod2_SVT_mat"
				    : "C_crossprod2_mat_SVT");
	if (rc != SVTGPU_OK)
		rglue_fail(rc, "svtgpu_crossprod");
}

/* --- .Call ENTRY POINT --- */
SEXP C_crossprod2_SVT_mat(SEXP x_dim, SEXP x_type, SEXP x_SVT, SEXP y,
			  SEXP transpose_y,
			  SEXP ans_type, SEXP ans_dimnames)
{
	int tr_y = LOGICAL(transpose_y)[0];

	SEXP y_dim = GET_DIM(y);
	if (LENGTH(x_dim) != 2 || LENGTH(y_dim) != 2)
		error("input objects must have 2 dimensions");
	int x_nrow = INTEGER(x_dim)[0];
	int x_ncol = INTEGER(x_dim)[1];
	int y_nrow = INTEGER(y_dim)[0];
	int y_ncol = INTEGER(y_dim)[1];
	if (x_nrow != (tr_y ? y_ncol : y_nrow))
		error("input objects are non-conformable");

	SEXPTYPE x_Rtype = get_and_check_input_Rtype(x_type, "x_type");
	if (x_Rtype != TYPEOF(y))
		error("SparseArray internal error in "
		      "C_crossprod2_SVT_mat():\n"
		      "    'x_Rtype != TYPEOF(y)' not supported yet");
	check_ans_type(ans_type, "C_crossprod2_SVT_mat");

	int ans_ncol = tr_y ? y_nrow : y_ncol;
	SEXP ans = PROTECT(new_double_matrix0(x_ncol, ans_ncol, ans_dimnames));
	/* x_SVT == NULL: all zeros, src/SparseMatrix_mult.c:389-390 */
	if (x_SVT != R_NilValue && XLENGTH(ans) != 0)
		run_crossprod(INTEGER(x_dim), x_Rtype, x_SVT, y, y_nrow,
			      y_ncol, tr_y, 1, REAL(ans));
	UNPROTECT(1);
	return ans;
}

/* --- .Call ENTRY POINT --- */
SEXP C_crossprod2_mat_SVT(SEXP x, SEXP y_dim, SEXP y_type, SEXP y_SVT,
			  SEXP transpose_x,
			  SEXP ans_type, SEXP ans_dimnames)
{
	int tr_x = LOGICAL(transpose_x)[0];

	SEXP x_dim = GET_DIM(x);
	if (LENGTH(x_dim) != 2 || LENGTH(y_dim) != 2)
		error("input objects must have 2 dimensions");
	int x_nrow = INTEGER(x_dim)[0];
	int x_ncol = INTEGER(x_dim)[1];
	int y_nrow = INTEGER(y_dim)[0];
	int y_ncol = INTEGER(y_dim)[1];
	if ((tr_x ? x_ncol : x_nrow) != y_nrow)
		error("input objects are non-conformable");

	SEXPTYPE y_Rtype = get_and_check_input_Rtype(y_type, "y_type");
	if (TYPEOF(x) != y_Rtype)
		error("input objects must have the same type() for now");
	check_ans_type(ans_type, "C_crossprod2_mat_SVT");

	int ans_nrow = tr_x ? x_nrow : x_ncol;
	SEXP ans = PROTECT(new_double_matrix0(ans_nrow, y_ncol, ans_dimnames));
	if (y_SVT != R_NilValue && XLENGTH(ans) != 0)
		run_crossprod(INTEGER(y_dim), y_Rtype, y_SVT, x, x_nrow,
			      x_ncol, tr_x, 0, REAL(ans));
	UNPROTECT(1);
	return ans;
}

/* --- .Call ENTRY POINT (extension) ---
 * ans = x %*% y with x an SVT_SparseMatrix and y an ordinary matrix of the
 * same type; same result as .crossprod2_SparseMatrix_matrix(t(x), y). */
SEXP C_matmul_SVT_mat(SEXP x_dim, SEXP x_type, SEXP x_SVT, SEXP y,
		      SEXP ans_dimnames)
{
	SEXP y_dim = GET_DIM(y);
	if (LENGTH(x_dim) != 2 || LENGTH(y_dim) != 2)
		error("input objects must have 2 dimensions");
	int x_nrow = INTEGER(x_dim)[0];
	int x_ncol = INTEGER(x_dim)[1];
	int y_nrow = INTEGER(y_dim)[0];
	int y_ncol = INTEGER(y_dim)[1];
	if (x_ncol != y_nrow)
		error("input objects are non-conformable");
	SEXPTYPE x_Rtype = get_and_check_input_Rtype(x_type, "x_type");
	if (x_Rtype != TYPEOF(y))
		error("SparseArray internal error in "
		      "C_matmul_SVT_mat():\n"
		      "    'x_Rtype != TYPEOF(y)' not supported yet");

	SEXP ans = PROTECT(new_double_matrix0(x_nrow, y_ncol, ans_dimnames));
	if (x_SVT != R_NilValue && XLENGTH(ans) != 0) {
		rglue_input in;
		rglue_acquire(x_SVT, INTEGER(x_dim), 2, x_Rtype, 1, 1, &in);
		int rc = svtgpu_matmul(in.m, DATAPTR(y), (int) x_Rtype, y_ncol,
				       REAL(ans));
		rglue_done(&in, "C_matmul_SVT_mat");
		if (rc != SVTGPU_OK)
			rglue_fail(rc, "svtgpu_matmul");
	}
	UNPROTECT(1);
	return ans;
}

/* --- .Call ENTRY POINT ---
 * crossprod(x, y), both SVT_SparseMatrix (R/SparseMatrix-mult.R:103-118;
 * reference src/SparseMatrix_mult.c:1037-1100). */
SEXP C_crossprod2_SVT_SVT(SEXP x_dim, SEXP x_type, SEXP x_SVT,
			  SEXP y_dim, SEXP y_type, SEXP y_SVT,
			  SEXP ans_type, SEXP ans_dimnames)
{
	if (LENGTH(x_dim) != 2 || LENGTH(y_dim) != 2)
		error("input objects must have 2 dimensions");
	int in_nrow = INTEGER(x_dim)[0];
	if (in_nrow != INTEGER(y_dim)[0])
		error("input SVT_SparseMatrix objects "
		      "are non-conformable");
	int x_ncol = INTEGER(x_dim)[1];
	int y_ncol = INTEGER(y_dim)[1];
	SEXPTYPE x_Rtype = get_and_check_input_Rtype(x_type, "x_type");
	SEXPTYPE y_Rtype = get_and_check_input_Rtype(y_type, "y_type");
	if (x_Rtype != y_Rtype)
		error("input SVT_SparseMatrix objects "
		      "must have the same type() for now");
	check_ans_type(ans_type, "C_crossprod2_SVT_SVT");

	SEXP ans = PROTECT(new_double_matrix0(x_ncol, y_ncol, ans_dimnames));
	if (XLENGTH(ans) == 0 || (x_SVT == R_NilValue && y_SVT == R_NilValue)) {
		UNPROTECT(1);
		return ans;
	}
	/* (a NULL x_SVT becomes an empty device matrix: svtgpu_crossprod_svt()
	   then applies the reference's "fictive matrix of zeros" rules,
	   crossprod2_mat0_SVT_*(), src/SparseMatrix_mult.c:558-611) */
	/* Order matters: everything that can raise an R error runs while no
	   per-call device matrix is held.  y == x (same SVT): one upload
	   serves both sides. */
	rglue_input inx, iny;
	int same = y_SVT == x_SVT && y_ncol == x_ncol;
	int y_is_handle = TYPEOF(y_SVT) == EXTPTRSXP;
	svt_leaf_index iy;
	if (!same && y_is_handle)
		rglue_acquire(y_SVT, INTEGER(y_dim), 2, y_Rtype, 1, 1, &iny);
	else if (!same)
		svt_index_leaves(y_SVT, INTEGER(y_dim), 2, y_Rtype, &iy);
	rglue_acquire(x_SVT, INTEGER(x_dim), 2, x_Rtype, 1, 1, &inx);
	if (same) {
		iny = inx;
	} else if (!y_is_handle) {
		memset(&iny, 0, sizeof(iny));
		int rcy = svt_upload_leaves(&iy, y_Rtype, 1, 1, &iny.m,
					    &iny.flatten_ms);
		iny.t_ready = rglue_now_ms();
		if (rcy != SVTGPU_OK) {
			rglue_done(&inx, "C_crossprod2_SVT_SVT");
			rglue_fail(rcy, "svt_upload_leaves");
		}
	}
	int rc = svtgpu_crossprod_svt(inx.m, iny.m, REAL(ans));
	if (!same)
		rglue_done(&iny, "C_crossprod2_SVT_SVT (y)");
	rglue_done(&inx, "C_crossprod2_SVT_SVT");
	if (rc != SVTGPU_OK)
		rglue_fail(rc, "svtgpu_crossprod_svt");
	UNPROTECT(1);
	return ans;
}

/* --- .Call ENTRY POINT ---
 * crossprod(x) (R/SparseMatrix-mult.R:120-133; reference
 * src/SparseMatrix_mult.c:1102-1140). */
SEXP C_crossprod1_SVT(SEXP x_dim, SEXP x_type, SEXP x_SVT,
		      SEXP ans_type, SEXP ans_dimnames)
{
	if (LENGTH(x_dim) != 2)
		error("'x' must have 2 dimensions");
	int x_ncol = INTEGER(x_dim)[1];
	SEXPTYPE x_Rtype = get_and_check_input_Rtype(x_type, "x_type");
	check_ans_type(ans_type, "C_crossprod1_SVT");
	SEXP ans = PROTECT(new_double_matrix0(x_ncol, x_ncol, ans_dimnames));
	if (x_SVT != R_NilValue && XLENGTH(ans) != 0 &&
	    INTEGER(x_dim)[0] != 0) {
		rglue_input in;
		rglue_acquire(x_SVT, INTEGER(x_dim), 2, x_Rtype, 1, 1, &in);
		int rc = svtgpu_crossprod_svt(in.m, in.m, REAL(ans));
		rglue_done(&in, "C_crossprod1_SVT");
		if (rc != SVTGPU_OK)
			rglue_fail(rc, "svtgpu_crossprod_svt");
	}
	UNPROTECT(1);
	return ans;
}
