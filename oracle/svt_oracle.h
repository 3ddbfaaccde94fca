/* TEST INFRASTRUCTURE -- CPU restatement ("port") of the reference algorithms
 * for the SVT hot path, on flat CSC arrays.  See svt_oracle.c. */
#ifndef SVT_ORACLE_H
#define SVT_ORACLE_H

#include <stdint.h>

/* A flattened SVT on the host.  vals == NULL: every leaf is lacunar;
 * lacunar != NULL: per-leaf flags (1 = nzvals is NULL, vals[] ignored). */
typedef struct svt_oracle_csc {
	int64_t nrow;            /* dim[0] */
	int64_t nleaf;           /* prod(dim[-1]) */
	const int64_t *leaf_ptr; /* nleaf + 1 */
	const int32_t *offs;
	const void *vals;        /* int32 or double, per val_type */
	int val_type;            /* 10 LGLSXP, 13 INTSXP, 14 REALSXP */
	const uint8_t *lacunar;
} svt_oracle_csc;

/* Return 0 on success, nonzero for an op/type the reference rejects. */
int svt_oracle_colstats(const svt_oracle_csc *x, int opcode, int narm,
			double center, int64_t group, void *out, int *warn);
int svt_oracle_rowstats(const svt_oracle_csc *x, int opcode, int narm,
			const double *center, void *out, int *warn);
int svt_oracle_crossprod(const svt_oracle_csc *x, const void *y,
			 int64_t y_nrow, int64_t y_ncol, int transpose_y,
			 int svt_on_left, double *ans);
/* t(x) as a CSC (caller frees the three arrays with free()). */
int svt_oracle_transpose(const svt_oracle_csc *x, int64_t **t_ptr,
			 int32_t **t_offs, void **t_vals);

/* rowsum(x, group) -> ngroup x nleaf; colsum(x, group) -> nrow x ngroup
 * (column-major, int32 for integer input, double for double input; `out` must
 * be zero-filled).  *overflow = 1 when the reference would warn "NAs produced
 * by integer overflow". */
int svt_oracle_rowsum(const svt_oracle_csc *x, const int32_t *group,
		      int32_t ngroup, int narm, void *out, int *overflow);
int svt_oracle_colsum(const svt_oracle_csc *x, const int32_t *group,
		      int32_t ngroup, int narm, void *out, int *overflow);

#endif
