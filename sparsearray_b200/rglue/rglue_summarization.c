/* .Call entry point C_summarize_SVT served by the GPU path: the whole-array
 * summaries behind sum(), prod(), mean(), var(), sd(), min(), max(), range(),
 * any(), all() and anyNA() of an SVT_SparseArray
 * (R/SparseArray-summarization.R:19-46).
 *
 * Same signature, argument checks, result type rules and warning as the
 * reference's entry point (src/SparseArray_summarization.c:112-142 and
 * res2nakedSEXP(), src/Rvector_summarization.c:1245-1301); the recursive tree
 * walk with its running SummarizeResult is replaced by: flatten the SVT ->
 * upload the values -> slice reduction + combine on the device
 * (svtgpu_summarize()).
 *
 * Not served (clean error): NaArray input, types other than
 * logical/integer/double, and "sum_X_X2" / "var2" / "sd2" (no R method
 * reaches them).
 */
#include "rglue_common.h"

#include "../csrc/svt_semantics.h"

#include <limits.h>

/* Round 'x' to nearest int (BACK_TO_INT, src/Rvector_summarization.c:1206) */
static int back_to_int(double x)
{
	return (int) (x >= 0 ? x + 0.5 : x - 0.5);
}

static SEXP scalar_logical(double v)
{
	/* a fresh vector: see the note on ScalarLogical2(),
	   src/Rvector_summarization.c:1184-1203 */
	SEXP ans = PROTECT(NEW_LOGICAL(1));
	LOGICAL(ans)[0] = ISNAN(v) ? NA_LOGICAL : v != 0.0;
	UNPROTECT(1);
	return ans;
}

static SEXP make_result(int opcode, SEXPTYPE in_Rtype, const double *out)
{
	if (opcode == SVTGPU_OP_ANYNA || opcode == SVTGPU_OP_ANY ||
	    opcode == SVTGPU_OP_ALL)
		return scalar_logical(out[0]);
	if (opcode == SVTGPU_OP_COUNTNAS) {
		if (out[0] > INT_MAX)
			return ScalarReal(out[0]);
		return ScalarInteger(back_to_int(out[0]));
	}
	if ((opcode == SVTGPU_OP_MIN || opcode == SVTGPU_OP_MAX) &&
	    in_Rtype != REALSXP)
		return ScalarInteger(ISNAN(out[0]) ? NA_INTEGER
						   : (int) out[0]);
	if (opcode == SVTGPU_OP_RANGE) {
		SEXP ans;
		if (in_Rtype == REALSXP) {
			ans = PROTECT(NEW_NUMERIC(2));
			REAL(ans)[0] = out[0];
			REAL(ans)[1] = out[1];
		} else {
			ans = PROTECT(NEW_INTEGER(2));
			for (int k = 0; k < 2; k++)
				INTEGER(ans)[k] = ISNAN(out[k]) ? NA_INTEGER
							       : (int) out[k];
		}
		UNPROTECT(1);
		return ans;
	}
	if ((opcode == SVTGPU_OP_SUM || opcode == SVTGPU_OP_PROD) &&
	    (in_Rtype == LGLSXP || in_Rtype == INTSXP)) {
		if (ISNAN(out[0]))
			return ScalarInteger(NA_INTEGER);
		if (out[0] < -INT_MAX || out[0] > INT_MAX)
			return ScalarReal(out[0]);
		return ScalarInteger(back_to_int(out[0]));
	}
	return ScalarReal(out[0]);
}

/* --- .Call ENTRY POINT --- */
SEXP C_summarize_SVT(SEXP x_dim, SEXP x_type, SEXP x_SVT,
		     SEXP x_na_background, SEXP op, SEXP na_rm, SEXP center)
{
	SEXPTYPE x_Rtype = rglue_get_and_check_Rtype(x_type,
					"C_summarize_SVT", "x_type");
	int x_has_NAbg = rglue_get_and_check_na_background(x_na_background,
					"C_summarize_SVT", "x_na_background");
	int opcode = rglue_get_summarize_opcode(op, x_Rtype);

	if (!(IS_LOGICAL(na_rm) && LENGTH(na_rm) == 1))
		error("'na.rm' must be TRUE or FALSE");
	int narm = LOGICAL(na_rm)[0];

	if (!IS_NUMERIC(center) || LENGTH(center) != 1)
		error("SparseArray internal error in "
		      "C_summarize_SVT():\n"
		      "    'center' must be a single number");

	if (x_has_NAbg)
		error("summarize: NaArray objects (na_background=TRUE) are "
		      "not supported by the SparseArray GPU path");
	if (x_Rtype != LGLSXP && x_Rtype != INTSXP && x_Rtype != REALSXP)
		error("summarize: SparseArray objects of type() \"%s\" are "
		      "not supported by the SparseArray GPU path",
		      type2char(x_Rtype));
	if (!svtgpu_summarize_supported(opcode, (int) x_Rtype))
		error("summarize: operation \"%s\" is not supported by the "
		      "SparseArray GPU path", CHAR(STRING_ELT(op, 0)));

	const int *dim = INTEGER(x_dim);
	int ndim = LENGTH(x_dim);
	double in_length = 1.0;
	for (int along = 0; along < ndim; along++)
		in_length *= dim[along];

	double out[2] = { 0.0, 0.0 };
	int warn = 0;
	if (in_length == 0.0 || ndim == 0) {
		/* nothing to summarise: the rules for an empty vector */
		SvtColPartial empty;
		svt_col_partial_init(&empty);
		int nres = opcode == SVTGPU_OP_RANGE ? 2 : 1;
		for (int k = 0; k < nres; k++) {
			int o = opcode != SVTGPU_OP_RANGE ? opcode
				: k == 0 ? SVTGPU_OP_MIN : SVTGPU_OP_MAX;
			double c = REAL(center)[0];
			if (svt_col_op_needs_center(o) && ISNAN(c))
				c = svt_col_mean(x_Rtype == REALSXP, narm, 0,
						 &empty);
			SvtScalar r = svt_col_finalize(o, x_Rtype == REALSXP,
						       narm, 0, c, &empty);
			if (svt_col_out_is_int(o, (int) x_Rtype))
				out[k] = r.i == NA_INTEGER ? NA_REAL
							   : (double) r.i;
			else
				out[k] = r.d;
			warn |= r.warn;
		}
	} else {
		/* a summary of all values never reads the row offsets */
		rglue_input in;
		rglue_acquire(x_SVT, dim, ndim, x_Rtype, 0, 1, &in);
		int rc = svtgpu_summarize(in.m, opcode, narm, REAL(center)[0],
					  out, &warn);
		rglue_done(&in, "C_summarize_SVT");
		if (rc != SVTGPU_OK)
			rglue_fail(rc, "svtgpu_summarize");
	}
	SEXP ans = PROTECT(make_result(opcode, x_Rtype, out));
	if (warn)
		warning("NAs introduced by coercion of "
			"infinite values to integers");
	UNPROTECT(1);
	return ans;
}
