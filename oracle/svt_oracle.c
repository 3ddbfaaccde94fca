/* TEST INFRASTRUCTURE -- not part of the product.
 *
 * CPU restatement ("port") of the algorithms Bioconductor/SparseArray runs for
 * SVT column/row statistics and SVT x dense crossprod, operating on a flat
 * CSC view of the SVT instead of R lists.  Scalar, sequential, same element
 * order and same accumulation order as the reference, so for any input it is
 * meant to return what the reference's C returns.
 *
 * Parity pinned: tests/test_oracle_vs_reference.py checks every function here
 * against the reference's own compiled C (oracle/_ref/libsvtref.so, built from
 * /root/reference/src by oracle/Makefile) on its test fixtures and on seeded
 * random inputs with NA/NaN/Inf injected; tests/test_golden.py checks both
 * against the known answers held in tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * link or call this file.
 */
#include "svt_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* opcodes: src/Rvector_summarization.h:12-32 */
enum { ANYNA = 1, COUNTNAS, ANY, ALL, MIN, MAX, RANGE, SUM, PROD, MEAN,
       X2SUM, SUM_X_X2, VAR1, VAR2, SD1, SD2 };

#define LGLSXP  10
#define INTSXP  13
#define REALSXP 14
#define NA_INT  INT_MIN

static double na_real(void)
{
	union { double d; uint64_t u; } x;
	x.u = 0x7FF00000000007A2ULL;  /* low word 1954 */
	return x.d;
}

static int is_na_real(double v)   /* R_IsNA() */
{
	union { double d; uint64_t u; } x;
	if (!isnan(v))
		return 0;
	x.d = v;
	return (uint32_t) x.u == 1954u;
}

static int is_nan_not_na(double v)   /* R_IsNaN() */
{
	return isnan(v) && !is_na_real(v);
}

/* ---- the summarisation engine (src/Rvector_summarization.c) ---- */

enum { NOT_SET = 1, IS_SET, BROKE };   /* OUTBUF_* :57-59 */

typedef struct {
	int64_t in_length, in_nzcount, in_nacount;
	int out_is_int;
	int status;
	int oi;          /* integer/logical output */
	double od;       /* double output */
	int one_zero;    /* postprocess_one_zero */
	int warn;
} res_t;

/* _init_SummarizeResult(), :97-165 */
static int res_init(res_t *r, int op, int type)
{
	memset(r, 0, sizeof(*r));
	r->status = IS_SET;
	switch (op) {
	    case ANYNA: case ANY: r->out_is_int = 1; r->oi = 0; return 0;
	    case COUNTNAS: r->od = 0.0; return 0;
	    case ALL: r->out_is_int = 1; r->oi = 1; r->one_zero = 1; return 0;
	    case SUM: case MEAN: case X2SUM: case VAR1: case SD1:
		r->od = 0.0; return 0;
	    case PROD: r->od = 1.0; r->one_zero = 1; return 0;
	    case MIN: case MAX:
		r->one_zero = 1;
		if (type == REALSXP) {
			r->od = op == MIN ? INFINITY : -INFINITY;
		} else {
			r->out_is_int = 1;
			r->status = NOT_SET;
		}
		return 0;
	}
	return 1;   /* RANGE, SUM_X_X2, VAR2, SD2: outside the path */
}

/* summarize_ints(), :827-866 and the loops it dispatches to */
static void eat_ints(res_t *r, const int *x, int64_t n, int op, int narm,
		     double center)
{
	int set_na = 0;
	for (int64_t i = 0; i < n; i++) {
		int v = x[i];
		if (op == ANYNA) {                        /* :177-187 */
			if (v == NA_INT) { r->oi = 1; r->status = BROKE; return; }
			continue;
		}
		if (op == COUNTNAS) {                     /* :226-235 */
			if (v == NA_INT) r->od++;
			continue;
		}
		if (v == NA_INT) {
			if (narm) { r->in_nacount++; continue; }
			if (op == ANY || op == ALL) { set_na = 1; continue; }
			if (r->out_is_int) r->oi = NA_INT;   /* :340-417 */
			else r->od = na_real();              /* :518-537 etc. */
			r->status = BROKE;
			return;
		}
		switch (op) {
		    case ANY:                             /* :261-286 */
			if (v != 0) { r->oi = 1; r->status = BROKE; return; }
			break;
		    case ALL:                             /* :291-316 */
			if (v == 0) { r->oi = 0; r->status = BROKE; return; }
			break;
		    case MIN:
			if (r->status == NOT_SET || v < r->oi) {
				r->oi = v; r->status = IS_SET;
			}
			break;
		    case MAX:
			if (r->status == NOT_SET || v > r->oi) {
				r->oi = v; r->status = IS_SET;
			}
			break;
		    case SUM: case MEAN: r->od += (double) v; break;
		    case PROD: r->od *= (double) v; break;
		    case X2SUM: case VAR1: case SD1: {    /* :622-642 */
			double delta = (double) v - center;
			r->od += delta * delta;
			break;
		    }
		}
	}
	if (set_na)
		r->oi = NA_INT;
}

/* summarize_doubles(), :868-904 and the loops it dispatches to */
static void eat_doubles(res_t *r, const double *x, int64_t n, int op,
			int narm, double center)
{
	double out = r->od;
	int out_ok = !is_nan_not_na(out);    /* out0_is_not_NaN */
	for (int64_t i = 0; i < n; i++) {
		double v = x[i];
		if (op == ANYNA) {                        /* :189-199 */
			if (isnan(v)) { r->oi = 1; r->status = BROKE; return; }
			continue;
		}
		if (op == COUNTNAS) {                     /* :237-246 */
			if (isnan(v)) out++;
			continue;
		}
		if (isnan(v)) {                           /* NA or NaN */
			if (narm) { r->in_nacount++; continue; }
			if (is_na_real(v)) {
				r->od = na_real();
				r->status = BROKE;
				return;
			}
			out = v;
			out_ok = 0;
			continue;
		}
		if (!out_ok)
			continue;
		switch (op) {
		    case MIN: if (v < out) out = v; break;   /* :365-392 */
		    case MAX: if (v > out) out = v; break;   /* :420-447 */
		    case SUM: case MEAN: out += v; break;    /* :540-567 */
		    case PROD: out *= v; break;              /* :592-619 */
		    case X2SUM: case VAR1: case SD1: {       /* :645-674 */
			double delta = v - center;
			out += delta * delta;
			break;
		    }
		}
	}
	r->od = out;
}

/* summarize_ones(), :742-825: a lacunar leaf of n implicit ones */
static void eat_ones(res_t *r, int64_t n, int op, int type, double center)
{
	if (n == 0)
		return;
	switch (op) {
	    case ANYNA: case COUNTNAS: case ALL: case PROD:
		return;
	    case ANY:
		r->oi = 1; r->status = BROKE;
		return;
	    case MIN:
		if (type != REALSXP) {
			if (r->status == NOT_SET || r->oi > 1) r->oi = 1;
		} else if (r->od > 1.0) {
			r->od = 1.0;
		}
		r->status = IS_SET;
		return;
	    case MAX:
		if (type != REALSXP) {
			if (r->status == NOT_SET || r->oi < 1) r->oi = 1;
		} else if (r->od < 1.0) {
			r->od = 1.0;
		}
		r->status = IS_SET;
		return;
	    case SUM: case MEAN:
		r->od += (double) n;
		return;
	    case X2SUM: case VAR1: case SD1: {
		double delta = 1.0 - center;
		r->od += delta * delta * n;
		return;
	    }
	}
}

/* summarize_leaf() + REC_summarize_SVT() over the `group` leaves of one
 * output cell: src/SparseArray_summarization.c:15-68 */
static void eat_segment(res_t *r, const svt_oracle_csc *x, int64_t leaf0,
			int64_t group, int op, int narm, double center)
{
	for (int64_t l = leaf0; l < leaf0 + group; l++) {
		int64_t start = x->leaf_ptr[l];
		int64_t nz = x->leaf_ptr[l + 1] - start;
		r->in_length += x->nrow;
		if (nz == 0)
			continue;              /* NULL leaf */
		r->in_nzcount += nz;
		if (x->vals == NULL || (x->lacunar != NULL && x->lacunar[l]))
			eat_ones(r, nz, op, x->val_type, center);
		else if (x->val_type == REALSXP)
			eat_doubles(r, (const double *) x->vals + start, nz,
				    op, narm, center);
		else
			eat_ints(r, (const int *) x->vals + start, nz,
				 op, narm, center);
		if (r->status == BROKE) {
			r->one_zero = 0;       /* :932-933, :994-995 */
			return;                /* bail out early, :64-65 */
		}
	}
}

/* _postprocess_SummarizeResult() with na_background = 0, :1078-1177 */
static void postprocess(res_t *r, int op, int type, int narm, double center)
{
	if (r->status == BROKE)
		return;
	int64_t zerocount = r->in_length - r->in_nzcount;
	if (op == COUNTNAS)
		return;
	int64_t effective_len = r->in_length;
	if (narm)
		effective_len -= r->in_nacount;
	if (zerocount != 0 && r->one_zero) {       /* summarize_one_zero() */
		if (type == REALSXP) {
			double zero = 0.0;
			eat_doubles(r, &zero, 1, op, narm, center);
		} else {
			int zero = 0;
			eat_ints(r, &zero, 1, op, narm, center);
		}
	}
	if (r->status == NOT_SET) {                /* :1108-1128 */
		r->oi = NA_INT;
		r->warn = 1;
		r->status = IS_SET;
		return;
	}
	switch (op) {
	    case MEAN:
		r->od /= (double) effective_len;
		return;
	    case X2SUM: case VAR1: case SD1:
		r->od += center * center * zerocount;
		if (op == X2SUM)
			return;
		if (effective_len <= 1) {
			r->od = na_real();
			return;
		}
		r->od /= (effective_len - 1.0);
		if (op == SD1)
			r->od = sqrt(r->od);
		return;
	}
}

/* _summarize_SVT(), src/SparseArray_summarization.c:70-109 */
static res_t summarize_segment(const svt_oracle_csc *x, int64_t leaf0,
			       int64_t group, int op, int narm, double center)
{
	res_t r;
	if ((op == X2SUM || op == VAR1 || op == SD1) && isnan(center)) {
		res_init(&r, MEAN, x->val_type);
		eat_segment(&r, x, leaf0, group, MEAN, narm, center);
		postprocess(&r, MEAN, x->val_type, narm, center);
		center = r.od;
	}
	res_init(&r, op, x->val_type);
	eat_segment(&r, x, leaf0, group, op, narm, center);
	postprocess(&r, op, x->val_type, narm, center);
	return r;
}

/* REC_colStats_SVT(), src/SparseArray_matrixStats.c:200-231 */
int svt_oracle_colstats(const svt_oracle_csc *x, int opcode, int narm,
			double center, int64_t group, void *out, int *warn)
{
	res_t probe;
	if (res_init(&probe, opcode, x->val_type))
		return 1;
	if ((opcode == ANY || opcode == ALL) && x->val_type == REALSXP)
		return 1;                      /* :69-72 */
	if (group < 1 || x->nleaf % group != 0)
		return 2;
	*warn = 0;
	int64_t nout = x->nleaf / group;
	for (int64_t s = 0; s < nout; s++) {
		res_t r = summarize_segment(x, s * group, group, opcode, narm,
					    center);
		if (r.warn)
			*warn = 1;
		if (r.out_is_int)
			((int *) out)[s] = r.oi;
		else
			((double *) out)[s] = r.od;
	}
	return 0;
}

/* ---- row statistics (src/SparseArray_matrixStats.c:303-1072) ---- */

/* update_out_with_int_{min,max}(), :303-347 */
static void row_upd_int(int x, int narm, int *out, int not_set, int is_min)
{
	if (narm) {
		if (x == NA_INT) return;
		if (*out == NA_INT) { *out = x; return; }
	} else {
		if (not_set || x == NA_INT) { *out = x; return; }
		if (*out == NA_INT) return;
	}
	if (is_min ? x < *out : x > *out)
		*out = x;
}

/* update_out_with_double_{min,max}(), :351-407 */
static void row_upd_double(double x, int narm, double *out, int not_set,
			   int is_min)
{
	if (narm) {
		if (isnan(x)) return;
		if (is_na_real(*out)) { *out = x; return; }
	} else {
		if (not_set || is_na_real(x)) { *out = x; return; }
		if (isnan(*out)) return;
		if (is_nan_not_na(x)) { *out = x; return; }
	}
	if (is_min ? x < *out : x > *out)
		*out = x;
}

int svt_oracle_rowstats(const svt_oracle_csc *x, int opcode, int narm,
			const double *center, void *out, int *warn)
{
	const int64_t nrow = x->nrow, nstrata = x->nleaf;
	const int is_double = x->val_type == REALSXP;
	const int out_is_int = opcode == ANYNA ||
		((opcode == MIN || opcode == MAX) && !is_double);
	int *oi = (int *) out;
	double *od = (double *) out;
	int64_t *nzcvg = NULL;
	*warn = 0;

	/* initialisation: SVT_row*() :856-1072 */
	switch (opcode) {
	    case ANYNA: case COUNTNAS: case SUM:
		for (int64_t i = 0; i < nrow; i++)
			if (out_is_int) oi[i] = 0; else od[i] = 0.0;
		break;
	    case X2SUM:
		for (int64_t i = 0; i < nrow; i++)
			od[i] = center == NULL ? 0.0
					: center[i] * center[i] * nstrata;
		break;
	    case MIN: case MAX:
		if (nstrata == 0) {                           /* :973-986 */
			for (int64_t i = 0; i < nrow; i++) {
				if (is_double)
					od[i] = opcode == MIN ? INFINITY
							      : -INFINITY;
				else
					oi[i] = NA_INT;
			}
			if (!is_double && nrow != 0)
				*warn = 1;
			return 0;
		}
		for (int64_t i = 0; i < nrow; i++) {
			if (is_double) od[i] = narm ? na_real() : 0.0;
			else           oi[i] = narm ? NA_INT : 0;
		}
		nzcvg = (int64_t *) calloc(nrow > 0 ? nrow : 1, sizeof(int64_t));
		break;
	    default:
		return 1;
	}
	if (nrow == 0)       /* C_rowStats_SVT() :1154-1157 */
		goto done;

	/* REC_rowStats_SVT(): leaves in order, :774-829 */
	for (int64_t l = 0; l < x->nleaf; l++) {
		int64_t start = x->leaf_ptr[l];
		int64_t nz = x->leaf_ptr[l + 1] - start;
		int lac = x->vals == NULL ||
			  (x->lacunar != NULL && x->lacunar[l]);
		const int *offs = x->offs + start;
		const int *iv = lac || is_double ? NULL
				: (const int *) x->vals + start;
		const double *dv = lac || !is_double ? NULL
				: (const double *) x->vals + start;
		for (int64_t k = 0; k < nz; k++) {
			int i = offs[k];
			switch (opcode) {
			    case ANYNA:                       /* :498-514 */
				if (lac) break;
				if (is_double ? isnan(dv[k]) : iv[k] == NA_INT)
					oi[i] = 1;
				break;
			    case COUNTNAS:                    /* :516-533 */
				if (lac) break;
				if (is_double ? isnan(dv[k]) : iv[k] == NA_INT)
					od[i]++;
				break;
			    case SUM:                         /* :599-613 */
				if (lac) { od[i] += 1.0; break; }
				if (is_double) {
					if (narm && isnan(dv[k])) break;
					od[i] += dv[k];
				} else {
					if (iv[k] == NA_INT) {
						if (narm) break;
						od[i] += na_real();
					} else {
						od[i] += (double) iv[k];
					}
				}
				break;
			    case X2SUM: {                     /* :636-696 */
				double c = center == NULL ? 0.0 : center[i];
				double v;
				if (lac) {
					double t = 1.0;
					if (center != NULL) t -= 2 * center[i];
					od[i] += t;
					break;
				}
				if (is_double) {
					v = dv[k];
					if (narm && isnan(v)) {
						od[i] -= c * c;
						break;
					}
				} else if (iv[k] == NA_INT) {
					if (narm) { od[i] -= c * c; break; }
					v = na_real();
				} else {
					v = (double) iv[k];
				}
				od[i] += v * (v - 2 * c);
				break;
			    }
			    case MIN: case MAX: {             /* :535-597 */
				int not_set = nzcvg[i]++ == 0;
				if (is_double)
					row_upd_double(lac ? 1.0 : dv[k], narm,
						od + i, not_set, opcode == MIN);
				else
					row_upd_int(lac ? 1 : iv[k], narm,
						oi + i, not_set, opcode == MIN);
				break;
			    }
			}
		}
	}

	/* postprocess_{int,double}_rowMinsMaxs(), :914-961 */
	if (opcode == MIN || opcode == MAX) {
		for (int64_t i = 0; i < nrow; i++) {
			if (nzcvg[i] < nstrata) {
				if (is_double)
					row_upd_double(0.0, narm, od + i,
						nzcvg[i] == 0, opcode == MIN);
				else
					row_upd_int(0, narm, oi + i,
						nzcvg[i] == 0, opcode == MIN);
			}
			if (is_double) {
				if (narm && is_na_real(od[i]))
					od[i] = opcode == MIN ? INFINITY
							      : -INFINITY;
			} else if (narm && oi[i] == NA_INT) {
				*warn = 1;
			}
		}
	}
done:
	free(nzcvg);
	return 0;
}

/* ---- crossprod (src/SparseVec_dotprod.c, src/SparseMatrix_mult.c) ---- */

/* _dotprod_doubleSV_finite_doubles() :28-43 / _dotprod_doubleSV_doubles()
 * :48-65 / _dotprod_doubles_zero() :116-126, chosen per dense column by
 * has_no_NaN_or_Inf() as in compute_dotprods2_with_double_Rcol() :193-207 */
static double dot_double(const svt_oracle_csc *x, int64_t l, const double *y,
			 int y_finite)
{
	int64_t start = x->leaf_ptr[l], nz = x->leaf_ptr[l + 1] - start;
	int lac = x->vals == NULL || (x->lacunar != NULL && x->lacunar[l]);
	const int *offs = x->offs + start;
	const double *v = lac ? NULL : (const double *) x->vals + start;
	double ans = 0.0;
	if (y_finite) {
		if (nz == 0)
			return 0.0;
		for (int64_t k = 0; k < nz; k++)
			ans += lac ? y[offs[k]] : v[k] * y[offs[k]];
		return ans;
	}
	int64_t k = 0;
	for (int64_t i = 0; i < x->nrow; i++) {
		double v1 = 0.0, v2 = y[i];
		if (is_na_real(v2))
			return na_real();
		if (k < nz && offs[k] == i) {
			v1 = lac ? 1.0 : v[k];
			if (is_na_real(v1))
				return na_real();
			k++;
		}
		ans += v1 * v2;
	}
	return ans;
}

/* _dotprod_intSV_noNA_ints() :73-92 / _dotprod_intSV_ints() :97-114 /
 * _dotprod_ints_zero() :128-138, chosen by has_no_NA() as in
 * compute_dotprods2_with_int_Rcol() :225-239 */
static double dot_int(const svt_oracle_csc *x, int64_t l, const int *y,
		      int y_no_na)
{
	int64_t start = x->leaf_ptr[l], nz = x->leaf_ptr[l + 1] - start;
	int lac = x->vals == NULL || (x->lacunar != NULL && x->lacunar[l]);
	const int *offs = x->offs + start;
	const int *v = lac ? NULL : (const int *) x->vals + start;
	double ans = 0.0;
	if (y_no_na) {
		for (int64_t k = 0; k < nz; k++) {
			if (lac) {
				ans += (double) y[offs[k]];
			} else {
				if (v[k] == NA_INT)
					return na_real();
				ans += (double) v[k] * y[offs[k]];
			}
		}
		return ans;
	}
	int64_t k = 0;
	for (int64_t i = 0; i < x->nrow; i++) {
		int v1 = 0, v2 = y[i];
		if (v2 == NA_INT)
			return na_real();
		if (k < nz && offs[k] == i) {
			v1 = lac ? 1 : v[k];
			if (v1 == NA_INT)
				return na_real();
			k++;
		}
		ans += (double) v1 * v2;
	}
	return ans;
}

/* crossprod2_SVT_mat_{double,int}() :385-431,483-514 and the mirror
 * crossprod2_mat_SVT_{double,int}() :435-479,518-547.  y is column-major
 * y_nrow x y_ncol of x's type.  ans: nleaf x K (svt_on_left) or K x nleaf,
 * column-major, pre-filled with zeros like _new_Rmatrix0(). */
int svt_oracle_crossprod(const svt_oracle_csc *x, const void *y,
			 int64_t y_nrow, int64_t y_ncol, int transpose_y,
			 int svt_on_left, double *ans)
{
	const int is_double = x->val_type == REALSXP;
	const int64_t K = transpose_y ? y_nrow : y_ncol;
	const int64_t in_nrow = transpose_y ? y_ncol : y_nrow;
	if (in_nrow != x->nrow)
		return 1;
	if (x->val_type != REALSXP && x->val_type != INTSXP)
		return 1;
	memset(ans, 0, sizeof(double) * (size_t) (x->nleaf * K));
	if (x->leaf_ptr[x->nleaf] == 0)
		return 0;       /* x_SVT == R_NilValue: :389-390 */
	size_t esz = is_double ? sizeof(double) : sizeof(int);
	char *colbuf = (char *) malloc(esz * (size_t) (in_nrow > 0 ? in_nrow : 1));
	for (int64_t k = 0; k < K; k++) {
		const void *col;
		if (transpose_y) {                            /* :411-421 */
			for (int64_t i = 0; i < in_nrow; i++)
				memcpy(colbuf + esz * i,
				       (const char *) y + esz * (k + i * y_nrow),
				       esz);
			col = colbuf;
		} else {
			col = (const char *) y + esz * (size_t) (k * y_nrow);
		}
		int clean = 1;
		for (int64_t i = 0; i < in_nrow && clean; i++)
			clean = is_double ? isfinite(((const double *) col)[i])
					  : ((const int *) col)[i] != NA_INT;
		for (int64_t l = 0; l < x->nleaf; l++) {
			double dp = is_double
				? dot_double(x, l, (const double *) col, clean)
				: dot_int(x, l, (const int *) col, clean);
			if (svt_on_left)
				ans[l + k * x->nleaf] = dp;
			else
				ans[k + l * K] = dp;
		}
	}
	free(colbuf);
	return 0;
}

/* transpose_2D_SVT(), src/SparseArray_aperm.c:348-401: count, allocate,
 * fill, walking the leaves in order (so every new leaf stays sorted). */
int svt_oracle_transpose(const svt_oracle_csc *x, int64_t **t_ptr,
			 int32_t **t_offs, void **t_vals)
{
	const int64_t nnz = x->leaf_ptr[x->nleaf];
	const int is_double = x->val_type == REALSXP;
	size_t esz = is_double ? sizeof(double) : sizeof(int);
	int64_t *ptr = (int64_t *) calloc((size_t) x->nrow + 1, sizeof(int64_t));
	int32_t *offs = (int32_t *) malloc(sizeof(int32_t) * (size_t) (nnz > 0 ? nnz : 1));
	char *vals = (char *) malloc(esz * (size_t) (nnz > 0 ? nnz : 1));
	int64_t *fill = (int64_t *) malloc(sizeof(int64_t) * (size_t) (x->nrow + 1));
	for (int64_t e = 0; e < nnz; e++)
		ptr[x->offs[e] + 1]++;
	for (int64_t i = 0; i < x->nrow; i++)
		ptr[i + 1] += ptr[i];
	memcpy(fill, ptr, sizeof(int64_t) * (size_t) (x->nrow + 1));
	for (int64_t l = 0; l < x->nleaf; l++) {
		int lac = x->vals == NULL ||
			  (x->lacunar != NULL && x->lacunar[l]);
		for (int64_t e = x->leaf_ptr[l]; e < x->leaf_ptr[l + 1]; e++) {
			int64_t dst = fill[x->offs[e]]++;
			offs[dst] = (int32_t) l;
			if (lac) {
				if (is_double) ((double *) vals)[dst] = 1.0;
				else           ((int *) vals)[dst] = 1;
			} else {
				memcpy(vals + esz * dst,
				       (const char *) x->vals + esz * e, esz);
			}
		}
	}
	free(fill);
	*t_ptr = ptr;
	*t_offs = offs;
	*t_vals = vals;
	return 0;
}

/* ------------------------------------------------------------------------
 * rowsum() / colsum()   (src/rowsum_methods.c)
 *
 * group[]: 1-based group of every row (rowsum) / column (colsum), NA_INTEGER
 * = the last group (:48-51, :217-220).  Integer sums go through S4Vectors'
 * safe_int_add() (NA in -> NA out; a result outside [-INT_MAX, INT_MAX] ->
 * NA + sticky overflow flag) for rowsum (:81) and through the double-typed
 * range check of add_sparse_vec_to_ints() (:172-199) for colsum; double sums
 * are plain `out += v` in storage order, NA / NaN skipped under na.rm.
 */
#include <limits.h>

static int oracle_safe_int_add(int x, int y, int *ovflow)
{
	if (x == INT_MIN || y == INT_MIN)
		return INT_MIN;
	if ((y > 0 && x > INT_MAX - y) || (y < 0 && x < -INT_MAX - y)) {
		*ovflow = 1;
		return INT_MIN;
	}
	return x + y;
}

/* out: ngroup x nleaf, column-major, zero-filled by the caller */
int svt_oracle_rowsum(const svt_oracle_csc *x, const int32_t *group,
		      int32_t ngroup, int narm, void *out, int *overflow)
{
	*overflow = 0;
	if (x->val_type != 13 && x->val_type != 14)
		return 1;   /* :313-318: integer and double only */
	for (int64_t l = 0; l < x->nleaf; l++) {
		const int lac = x->vals == NULL ||
				(x->lacunar != NULL && x->lacunar[l]);
		for (int64_t k = x->leaf_ptr[l]; k < x->leaf_ptr[l + 1]; k++) {
			int g = group[x->offs[k]];
			if (g == INT_MIN)
				g = ngroup;
			g--;
			if (x->val_type == 14) {
				double *o = (double *) out + l * ngroup;
				double v = 1.0;
				if (!lac) {
					v = ((const double *) x->vals)[k];
					if (narm && isnan(v))
						continue;
				}
				/* same operand order as the reference build
				   (`addsd v, [out]`): of two NaNs the second
				   operand of `+=`, i.e. v, gives the payload */
				o[g] = v + o[g];
			} else {
				int *o = (int *) out + l * ngroup;
				int v = 1;
				if (!lac) {
					v = ((const int *) x->vals)[k];
					if (narm && v == INT_MIN)
						continue;
				}
				o[g] = oracle_safe_int_add(o[g], v, overflow);
			}
		}
	}
	return 0;
}

/* out: nrow x ngroup, column-major, zero-filled by the caller */
int svt_oracle_colsum(const svt_oracle_csc *x, const int32_t *group,
		      int32_t ngroup, int narm, void *out, int *overflow)
{
	*overflow = 0;
	if (x->val_type != 13 && x->val_type != 14)
		return 1;
	for (int64_t l = 0; l < x->nleaf; l++) {
		const int lac = x->vals == NULL ||
				(x->lacunar != NULL && x->lacunar[l]);
		int g = group[l];
		if (g == INT_MIN)
			g = ngroup;
		g--;
		for (int64_t k = x->leaf_ptr[l]; k < x->leaf_ptr[l + 1]; k++) {
			const int64_t at = (int64_t) g * x->nrow + x->offs[k];
			if (x->val_type == 14) {
				double *o = (double *) out + at;
				double v = 1.0;
				if (!lac) {
					v = ((const double *) x->vals)[k];
					if (narm && isnan(v))
						continue;
				}
				*o = v + *o;
			} else {
				int *o = (int *) out + at;
				if (*o == INT_MIN)
					continue;
				int v = 1;
				if (!lac) {
					v = ((const int *) x->vals)[k];
					if (v == INT_MIN) {
						if (!narm)
							*o = INT_MIN;
						continue;
					}
				}
				double y = (double) *o + v;
				if (-INT_MAX <= y && y <= INT_MAX) {
					*o = (int) y;
				} else {
					*overflow = 1;
					*o = INT_MIN;
				}
			}
		}
	}
	return 0;
}
