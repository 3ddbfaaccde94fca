/* svt_semantics.h -- result composition rules of the reference, shared by the
 * CUDA kernels (device) and by tests/semantics_host.cpp (host build of the
 * very same functions, so the NA/NaN/zero-background logic can be checked on
 * a machine without a GPU).  No loops over data here: the kernels reduce a
 * leaf (or a row) to a small "partial" and these functions turn a partial
 * into the value the reference would have produced.
 *
 * Reference: src/Rvector_summarization.c (per-type loops :177-734, lacunar
 * :742-825, post-processing :1078-1177), src/SparseArray_summarization.c
 * :70-109 (two-pass variance), src/SparseArray_matrixStats.c (:303-430 row
 * update rules, :914-961 row min/max post-processing, :1044-1072 centered
 * X2 sums), src/SparseVec_dotprod.c:28-138 and src/SparseMatrix_mult.c
 * :193-239 (dot-product NA rules).
 */
#ifndef SVT_SEMANTICS_H
#define SVT_SEMANTICS_H

#include <stdint.h>
#include <string.h>
#include <math.h>
#ifndef __cplusplus
#include <stdbool.h>
#endif

#include "../../include/svtgpu.h"

#ifdef __CUDACC__
#define SVT_HD __host__ __device__ __forceinline__
#else
#define SVT_HD static inline
#endif

#define SVT_NA_INT INT32_MIN

SVT_HD uint64_t svt_d2u(double x)
{
#ifdef __CUDA_ARCH__
	return (uint64_t) __double_as_longlong(x);
#else
	uint64_t u;
	memcpy(&u, &x, sizeof(u));
	return u;
#endif
}

SVT_HD double svt_u2d(uint64_t u)
{
#ifdef __CUDA_ARCH__
	return __longlong_as_double((long long) u);
#else
	double x;
	memcpy(&x, &u, sizeof(x));
	return x;
#endif
}

/* R's NA_real_ (arithmetic.c: low word 1954) and the canonical quiet NaN. */
SVT_HD double svt_na_real(void) { return svt_u2d(0x7FF00000000007A2ULL); }
SVT_HD double svt_nan(void)     { return svt_u2d(0x7FF8000000000000ULL); }
SVT_HD double svt_posinf(void)  { return svt_u2d(0x7FF0000000000000ULL); }
SVT_HD double svt_neginf(void)  { return svt_u2d(0xFFF0000000000000ULL); }

SVT_HD bool svt_isnan(double x)
{
	return (svt_d2u(x) & 0x7FFFFFFFFFFFFFFFULL) > 0x7FF0000000000000ULL;
}

/* R_IsNA(): a NaN whose low word is 1954. */
SVT_HD bool svt_is_na_real(double x)
{
	return svt_isnan(x) && (uint32_t) svt_d2u(x) == 1954u;
}

SVT_HD bool svt_isfinite(double x)
{
	return (svt_d2u(x) & 0x7FF0000000000000ULL) != 0x7FF0000000000000ULL;
}

/* Never let an accidental payload decide between NA and NaN on output. */
SVT_HD double svt_clean_nan(double x)
{
	return svt_isnan(x) ? svt_nan() : x;
}

/* ------------------------------------------------------------------------
 * Column statistics
 */

/* What a kernel knows about one summarised segment (a leaf, or `group`
 * consecutive leaves) once its stored values have been reduced. */
typedef struct SvtColPartial {
	int64_t nz;      /* stored values (in_nzcount) */
	int64_t n_na;    /* int: == NA_INTEGER; double: R_IsNA() */
	int64_t n_nan;   /* double only: NaN that is not NA */
	int64_t n_zero;  /* stored values equal to 0 (ANY/ALL only) */
	double sum;      /* sum of regular (non-NA, non-NaN) values */
	double sum2;     /* sum of (x - center)^2 over regular values */
	double prod;     /* product of regular values */
	double vmin;     /* min / max over regular values, +Inf / -Inf if none */
	double vmax;
} SvtColPartial;

SVT_HD void svt_col_partial_init(SvtColPartial *p)
{
	p->nz = p->n_na = p->n_nan = p->n_zero = 0;
	p->sum = p->sum2 = 0.0;
	p->prod = 1.0;
	p->vmin = svt_posinf();
	p->vmax = svt_neginf();
}

/* Partial of a lacunar segment: nz implicit ones
 * (summarize_ones(), src/Rvector_summarization.c:742-825). */
SVT_HD void svt_col_partial_ones(SvtColPartial *p, int64_t nz, double center)
{
	svt_col_partial_init(p);
	p->nz = nz;
	if (nz == 0)
		return;
	p->sum = (double) nz;
	double delta = 1.0 - center;
	p->sum2 = delta * delta * (double) nz;
	p->vmin = p->vmax = 1.0;
}

SVT_HD int svt_col_out_is_int(int opcode, int val_type)
{
	if (opcode == SVTGPU_OP_ANYNA || opcode == SVTGPU_OP_ANY ||
	    opcode == SVTGPU_OP_ALL)
		return 1;
	if ((opcode == SVTGPU_OP_MIN || opcode == SVTGPU_OP_MAX) &&
	    val_type != SVTGPU_DOUBLE)
		return 1;
	return 0;
}

SVT_HD int svt_col_op_supported(int opcode, int val_type)
{
	switch (opcode) {
	    case SVTGPU_OP_ANYNA: case SVTGPU_OP_COUNTNAS:
	    case SVTGPU_OP_MIN: case SVTGPU_OP_MAX:
	    case SVTGPU_OP_SUM: case SVTGPU_OP_PROD: case SVTGPU_OP_MEAN:
	    case SVTGPU_OP_CENTERED_X2_SUM:
	    case SVTGPU_OP_VAR1: case SVTGPU_OP_SD1:
		return 1;
	    case SVTGPU_OP_ANY: case SVTGPU_OP_ALL:
		return val_type != SVTGPU_DOUBLE;
	}
	return 0;
}

SVT_HD int svt_col_op_needs_center(int opcode)
{
	return opcode == SVTGPU_OP_CENTERED_X2_SUM ||
	       opcode == SVTGPU_OP_VAR1 || opcode == SVTGPU_OP_SD1;
}

typedef struct SvtScalar {
	double d;   /* result when the output type is double */
	int32_t i;  /* result when the output type is int32/logical */
	int warn;
} SvtScalar;

/* mean of the segment as computed by replace_SummarizeOp_center_with_mean()
 * (src/SparseArray_summarization.c:70-87): MEAN_OPCODE with the same na.rm. */
SVT_HD double svt_col_mean(int is_double, int narm, int64_t in_length,
			   const SvtColPartial *p)
{
	if (!narm && p->n_na > 0)
		return svt_na_real();   /* breaking value, returned as is */
	double s = (!narm && p->n_nan > 0) ? svt_nan() : p->sum;
	int64_t nacount = narm ? p->n_na + p->n_nan : 0;
	(void) is_double;
	return s / (double) (in_length - nacount);
}

/* Turn the partial of a segment of virtual length `in_length` into the value
 * _summarize_SVT() + _postprocess_SummarizeResult() produce (na_background
 * FALSE).  For CENTERED_X2_SUM/VAR1/SD1 `center` must already be the mean when
 * the caller's center was NA/NaN, and p->sum2 must be relative to it. */
SVT_HD SvtScalar svt_col_finalize(int opcode, int is_double, int narm,
				  int64_t in_length, double center,
				  const SvtColPartial *p)
{
	SvtScalar r;
	r.d = 0.0; r.i = 0; r.warn = 0;
	const int64_t zerocount = in_length - p->nz;
	const int64_t n_nanish = p->n_na + p->n_nan;
	const int64_t n_reg = p->nz - n_nanish;

	if (opcode == SVTGPU_OP_ANYNA) {
		r.i = n_nanish > 0;
		return r;
	}
	if (opcode == SVTGPU_OP_COUNTNAS) {
		r.d = (double) n_nanish;
		return r;
	}
	if (opcode == SVTGPU_OP_ANY) {
		/* any_ints(), :261-286: a nonzero non-NA value wins, then NA. */
		if (n_reg - p->n_zero > 0)
			r.i = 1;
		else if (!narm && p->n_na > 0)
			r.i = SVT_NA_INT;
		else
			r.i = 0;
		return r;
	}
	if (opcode == SVTGPU_OP_ALL) {
		/* all_ints(), :291-316, then one background zero, :1100-1106 */
		if (p->n_zero > 0 || zerocount > 0)
			r.i = 0;
		else if (!narm && p->n_na > 0)
			r.i = SVT_NA_INT;
		else
			r.i = 1;
		return r;
	}

	/* Every remaining op bails out with NA on the first NA when !na.rm
	   (OUTBUF_IS_SET_WITH_BREAKING_VALUE: post-processing is skipped). */
	const int broke = !narm && p->n_na > 0;
	/* double input, !na.rm: a NaN sticks but does not break (:548-561). */
	const int stuck_nan = !narm && p->n_nan > 0;
	const int64_t nacount = narm ? n_nanish : 0;
	const int64_t effective_len = in_length - nacount;

	switch (opcode) {
	    case SVTGPU_OP_MIN: case SVTGPU_OP_MAX: {
		const int is_min = opcode == SVTGPU_OP_MIN;
		if (!is_double) {
			if (broke) {
				r.i = SVT_NA_INT;
				return r;
			}
			int have = n_reg > 0;
			double v = is_min ? p->vmin : p->vmax;
			if (zerocount > 0) {
				if (!have || (is_min ? 0.0 < v : 0.0 > v))
					v = 0.0;
				have = 1;
			}
			if (!have) {
				/* empty or all-NA with na.rm: :1108-1127 */
				r.i = SVT_NA_INT;
				r.warn = 1;
				return r;
			}
			r.i = (int32_t) v;
			return r;
		}
		if (broke) {
			r.d = svt_na_real();
			return r;
		}
		if (stuck_nan) {
			r.d = svt_nan();  /* the background zero cannot undo it */
			return r;
		}
		double v = is_min ? p->vmin : p->vmax;
		if (zerocount > 0 && (is_min ? 0.0 < v : 0.0 > v))
			v = 0.0;
		r.d = v;
		return r;
	    }
	    case SVTGPU_OP_SUM: case SVTGPU_OP_MEAN: {
		if (broke) {
			r.d = svt_na_real();
			return r;
		}
		double v = stuck_nan ? svt_nan() : p->sum;
		if (opcode == SVTGPU_OP_MEAN)
			v = v / (double) effective_len;
		r.d = svt_clean_nan(v);
		return r;
	    }
	    case SVTGPU_OP_PROD: {
		if (broke) {
			r.d = svt_na_real();
			return r;
		}
		double v = stuck_nan ? svt_nan() : p->prod;
		if (zerocount > 0)
			v = v * 0.0;  /* Inf * 0 -> NaN, as prod_*(&zero, 1) */
		r.d = svt_clean_nan(v);
		return r;
	    }
	    case SVTGPU_OP_CENTERED_X2_SUM:
	    case SVTGPU_OP_VAR1: case SVTGPU_OP_SD1: {
		if (broke) {
			r.d = svt_na_real();
			return r;
		}
		double v = stuck_nan ? svt_nan() : p->sum2;
		v += center * center * (double) zerocount;      /* :1145-1147 */
		if (opcode == SVTGPU_OP_CENTERED_X2_SUM) {
			r.d = svt_clean_nan(v);
			return r;
		}
		if (effective_len <= 1) {
			r.d = svt_na_real();
			return r;
		}
		v /= ((double) effective_len - 1.0);
		if (opcode == SVTGPU_OP_SD1)
			v = sqrt(v);
		r.d = svt_clean_nan(v);
		return r;
	    }
	}
	r.d = svt_nan();
	return r;
}

/* ------------------------------------------------------------------------
 * Row statistics: per-row state slots (arrays of nrow doubles)
 */

/* slot indices */
#define SVT_ROW_SLOT_SUM    0  /* sum of regular values */
#define SVT_ROW_SLOT_NA     1  /* # NA */
#define SVT_ROW_SLOT_NAN    2  /* # NaN (double input) */
#define SVT_ROW_SLOT_SUM2   3  /* sum of squares of regular values */
/* MIN/MAX use: 0 = coverage (# leaves with a stored value in the row),
   1 = #NA, 2 = #NaN, 3 = running min or max over regular values. */
#define SVT_ROW_SLOT_CVG    0
#define SVT_ROW_SLOT_EXT    3
/* SUM / CENTERED_X2_SUM also carry, MAX-combined, the (global) index of the
   LAST leaf that put an NA / a NaN into the row (-Inf = none, or not tracked
   by the kernel that built the state).  The reference adds the entries of a
   row in leaf order with a plain `*out += x`
   (src/SparseArray_matrixStats.c:409-430); when both operands of that
   addition are NaNs the hardware keeps the payload of the first operand, and
   the reference as compiled (gcc, x86-64: `addsd x, [out]`) has x first --
   so of several NA / NaN entries in one row the last one decides whether the
   row sum is NA_real_ or NaN.  (+Inf and -Inf in one row make a fresh NaN,
   which any NA / NaN entry, earlier or later, overrides the same way.) */
#define SVT_ROW_SLOT_LAST_NA     4
#define SVT_ROW_SLOT_LAST_NAN    5

SVT_HD int svt_row_op_supported(int opcode)
{
	return opcode == SVTGPU_OP_ANYNA || opcode == SVTGPU_OP_COUNTNAS ||
	       opcode == SVTGPU_OP_MIN || opcode == SVTGPU_OP_MAX ||
	       opcode == SVTGPU_OP_SUM ||
	       opcode == SVTGPU_OP_CENTERED_X2_SUM;
}

/* number of SUM-combined and MIN/MAX-combined slots of an op's state */
SVT_HD void svt_row_state_layout(int opcode, int *n_sum, int *n_ext)
{
	switch (opcode) {
	    case SVTGPU_OP_ANYNA: case SVTGPU_OP_COUNTNAS:
		*n_sum = 3; *n_ext = 0; return;
	    case SVTGPU_OP_SUM:
	    case SVTGPU_OP_CENTERED_X2_SUM:
		*n_sum = 4; *n_ext = 2; return;
	    case SVTGPU_OP_MIN: case SVTGPU_OP_MAX:
		*n_sum = 3; *n_ext = 1; return;
	}
	*n_sum = 0; *n_ext = 0;
}

/* Which NaN a row sum ends up with under the reference's sequential
 * `*out += x` (na.rm = FALSE): 0 = none (the sum of the regular values
 * stands), 1 = NA_real_, 2 = NaN: the kind of the last NA / NaN entry in leaf
 * order.  When the positions were not tracked (-Inf in the slot although the
 * count is positive) NA wins. */
SVT_HD int svt_row_sum_kind(const double *s, int64_t stride)
{
	const double ninf = svt_neginf();
	const double n_na = s[SVT_ROW_SLOT_NA * stride];
	const double n_nan = s[SVT_ROW_SLOT_NAN * stride];
	if (n_na <= 0.0 && n_nan <= 0.0)
		return 0;
	const double l_na = s[SVT_ROW_SLOT_LAST_NA * stride];
	const double l_nan = s[SVT_ROW_SLOT_LAST_NAN * stride];
	if ((n_na > 0.0 && !(l_na > ninf)) || (n_nan > 0.0 && !(l_nan > ninf)))
		return n_na > 0.0 ? 1 : 2;
	if (n_na <= 0.0)
		return 2;
	if (n_nan <= 0.0)
		return 1;
	return l_na > l_nan ? 1 : 2;
}

/* One row's result from its combined state.  `s` points at slot 0 of the row,
 * consecutive slots are `stride` doubles apart. */
SVT_HD SvtScalar svt_row_finalize(int opcode, int is_double, int narm,
				  int64_t nstrata, int have_center,
				  double center, const double *s,
				  int64_t stride)
{
	SvtScalar r;
	r.d = 0.0; r.i = 0; r.warn = 0;
	const double n_na = s[SVT_ROW_SLOT_NA * stride];
	const double n_nan = s[SVT_ROW_SLOT_NAN * stride];

	switch (opcode) {
	    case SVTGPU_OP_ANYNA:
		r.i = (n_na + n_nan) > 0.0;
		return r;
	    case SVTGPU_OP_COUNTNAS:
		r.d = n_na + n_nan;
		return r;
	    case SVTGPU_OP_SUM: {
		/* update_out_with_{int,double}_sum(), :409-430: a plain
		   running `out += x`, so NA/NaN survive unless na.rm. */
		const int kind = narm ? 0 : svt_row_sum_kind(s, stride);
		if (kind == 1)
			r.d = svt_na_real();
		else if (kind == 2)
			r.d = svt_nan();
		else
			r.d = svt_clean_nan(s[SVT_ROW_SLOT_SUM * stride]);
		return r;
	    }
	    case SVTGPU_OP_CENTERED_X2_SUM: {
		/* SVT_rowCenteredX2Sum(), :1044-1072, with the per-nonzero
		   term x*(x - 2c) (:693) summed as sum(x^2) - 2c*sum(x). */
		/* the running value starts at c^2 * ncol (:1052-1058) */
		if (have_center && svt_isnan(center)) {
			r.d = svt_is_na_real(center) ? svt_na_real() : svt_nan();
			return r;
		}
		const int kind = narm ? 0 : svt_row_sum_kind(s, stride);
		if (kind == 1) {
			r.d = svt_na_real();
			return r;
		}
		if (kind == 2) {
			r.d = svt_nan();
			return r;
		}
		double c = have_center ? center : 0.0;
		double sx = s[SVT_ROW_SLOT_SUM * stride];
		double sx2 = s[SVT_ROW_SLOT_SUM2 * stride];
		double v;
		if (!have_center) {
			v = sx2;        /* c == 0: no 0 * Inf from the cross term */
		} else if (sx2 == svt_posinf() && svt_isfinite(c)) {
			/* an infinite x contributes x * (x - 2c) = +Inf whatever
			   its sign; Inf - Inf in the expanded form must not
			   turn that into NaN */
			v = sx2;
		} else {
			v = c * c * (double) nstrata + (sx2 - 2.0 * c * sx);
		}
		if (narm)
			v -= (n_na + n_nan) * (c * c);   /* :665-669,681-684 */
		r.d = svt_clean_nan(v);
		return r;
	    }
	    case SVTGPU_OP_MIN: case SVTGPU_OP_MAX: {
		const int is_min = opcode == SVTGPU_OP_MIN;
		const double cvg = s[SVT_ROW_SLOT_CVG * stride];
		const double n_reg = cvg - n_na - n_nan;
		double v = s[SVT_ROW_SLOT_EXT * stride];
		int have = n_reg > 0.0;
		if (nstrata == 0) {             /* SVT_rowMinsMaxs(), :973-986 */
			if (is_double) {
				r.d = is_min ? svt_posinf() : svt_neginf();
			} else {
				r.i = SVT_NA_INT;
				r.warn = 1;
			}
			return r;
		}
		if (!narm && n_na > 0.0) {
			if (is_double) r.d = svt_na_real();
			else           r.i = SVT_NA_INT;
			return r;
		}
		if (!narm && n_nan > 0.0) {
			r.d = svt_nan();
			return r;
		}
		if (cvg < (double) nstrata) {   /* fold one background zero */
			if (!have || (is_min ? 0.0 < v : 0.0 > v))
				v = 0.0;
			have = 1;
		}
		if (!have) {
			/* na.rm and nothing but NAs: :931-933, :955-956 */
			if (is_double) {
				r.d = is_min ? svt_posinf() : svt_neginf();
			} else {
				r.i = SVT_NA_INT;
				r.warn = 1;
			}
			return r;
		}
		if (is_double) r.d = v;
		else           r.i = (int32_t) v;
		return r;
	    }
	}
	r.d = svt_nan();
	return r;
}

/* rowMeans()/rowVars(center=NULL) as composed in R
 * (R/SparseArray-matrixStats.R:300-310,511-517,645-661) from one state
 * {sum x, #NA, #NaN, sum x^2}. */
SVT_HD void svt_row_moments(int narm, int64_t nstrata, const double *s,
			    int64_t stride, double *mean, double *var)
{
	const double n_na = s[SVT_ROW_SLOT_NA * stride];
	const double n_nan = s[SVT_ROW_SLOT_NAN * stride];
	const double sx = s[SVT_ROW_SLOT_SUM * stride];
	const double sx2 = s[SVT_ROW_SLOT_SUM2 * stride];
	const double nvals = (double) nstrata - (narm ? n_na + n_nan : 0.0);
	double sums;
	const int kind = narm ? 0 : svt_row_sum_kind(s, stride);
	if (kind == 1)      sums = svt_na_real();
	else if (kind == 2) sums = svt_nan();
	else                sums = sx;
	const int is_na = svt_is_na_real(sums);
	double c = sums / nvals;
	*mean = is_na ? svt_na_real() : svt_clean_nan(c);
	double x2 = c * c * (double) nstrata + (sx2 - 2.0 * c * sx);
	if (narm)
		x2 -= (n_na + n_nan) * (c * c);
	double v = x2 / (nvals - 1.0);
	*var = is_na ? svt_na_real() : svt_clean_nan(v);
}

/* ------------------------------------------------------------------------
 * SVT x dense dot products
 */

/* What is known about one column of the dense operand. */
typedef struct SvtDenseColInfo {
	int32_t n_nonfinite;  /* double: !R_FINITE; int: == NA_INTEGER;
				 SVT_COL_FICTIVE_ZERO: the column belongs to a
				 fictive all-zero operand (a NULL SVT) */
	int32_t n_na;         /* double: R_IsNA;    int: == NA_INTEGER */
} SvtDenseColInfo;
#define SVT_COL_FICTIVE_ZERO (-1)

/* Result of dot(leaf, y[,k]) given the plain gathered sum `s`
 * (sum of v * y[off] over the leaf's stored values), what is known about the
 * leaf's NA / NaN entries, and how many of its nonzeros hit non-finite entries
 * of the column.
 *   leaf_flag bit 0: the leaf holds an NA
 *             bit 1: the first NA-or-NaN entry of the leaf is a NaN
 * int: _dotprod_intSV_noNA_ints()/_dotprod_intSV_ints()/_dotprod_ints_zero();
 * double: _dotprod_doubleSV_finite_doubles()/_dotprod_doubleSV_doubles()/
 * _dotprod_doubles_zero() (src/SparseVec_dotprod.c:28-138), selected per
 * dense column as in src/SparseMatrix_mult.c:193-239.  The finite-column loop
 * is `ans += v * y` on a register accumulator (:36-41): of several NaNs the
 * first one's payload survives, so NA_real_ comes out only when the NA is the
 * leaf's first NA/NaN entry; the dense walk (:52-63) returns NA_real_ as soon
 * as it meets any NA. */
#define SVT_LEAF_HAS_NA     1
#define SVT_LEAF_NAN_FIRST  2

SVT_HD double svt_dot_finalize(int is_double, double s, int leaf_flag,
			       int64_t hits_nonfinite, SvtDenseColInfo ci)
{
	const int leaf_has_na = leaf_flag & SVT_LEAF_HAS_NA;
	if (!is_double) {
		if (ci.n_na > 0 || leaf_has_na)
			return svt_na_real();
		return s;
	}
	if (ci.n_nonfinite == SVT_COL_FICTIVE_ZERO) {
		/* the other operand is a NULL SVT: _dotprod_doubleSV_zero() /
		   _dotprod_doubles_zero() (src/SparseVec_dotprod.c:116-147)
		   return NA as soon as ANY entry is NA, wherever it stands;
		   otherwise the sum of v * 0 (NaN when the leaf holds a NaN or
		   an infinity) */
		if (leaf_has_na)
			return svt_na_real();
		return svt_clean_nan(s);
	}
	if (ci.n_nonfinite == 0) {
		/* fast path: NA/NaN leaf values propagate arithmetically */
		if (leaf_has_na && !(leaf_flag & SVT_LEAF_NAN_FIRST))
			return svt_na_real();
		return svt_clean_nan(s);
	}
	/* dense walk over all rows: NA wins, then any 0 * non-finite is NaN */
	if (ci.n_na > 0 || leaf_has_na)
		return svt_na_real();
	if (hits_nonfinite < (int64_t) ci.n_nonfinite)
		return svt_nan();
	return svt_clean_nan(s);
}

#endif  /* SVT_SEMANTICS_H */
