/* the narrowing copies below are plain loops meant to be vectorised */
#pragma GCC optimize("O3")
#include "svt_flatten.h"

#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef _OPENMP
#undef match
#include <omp.h>
#endif

/* Validates a leaf the way unzip_leaf()/toSparseVec() do and records its
 * payload pointers.  Returns 0, or an error code (no R API that can longjmp in
 * here: the matrix case runs this from OpenMP threads, as the reference does
 * with VECTOR_ELT() in its own parallel loops, src/SparseMatrix_mult.c:134). */
static const char *const leaf_errors[] = {
	NULL,
	"invalid SVT leaf",
	"TYPEOF(nzvals) != Rtype",
	"invalid SVT leaf ('nzvals' and 'nzoffs' are not parallel)",
};

static int index_leaf(SEXP leaf, SEXPTYPE Rtype, int64_t l,
		      svt_leaf_index *ix, int *is_lacunar)
{
	*is_lacunar = 0;
	if (!isVectorList(leaf) || LENGTH(leaf) < 2)
		return 1;
	SEXP nzvals = VECTOR_ELT(leaf, 0);
	SEXP nzoffs = VECTOR_ELT(leaf, 1);
	if (!IS_INTEGER(nzoffs))
		return 1;
	R_xlen_t nzcount = XLENGTH(nzoffs);
	if (nzcount == 0 || nzcount > INT_MAX)
		return 1;
	ix->leaf_ptr[l + 1] = nzcount;
	ix->offs[l] = INTEGER(nzoffs);
	if (nzvals == R_NilValue) {
		ix->vals[l] = NULL;
		*is_lacunar = 1;
		return 0;
	}
	if (TYPEOF(nzvals) != Rtype)
		return 2;
	if (XLENGTH(nzvals) != nzcount)
		return 3;
	ix->vals[l] = DATAPTR(nzvals);
	return 0;
}

static void leaf_error(int code)
{
	error("SparseArray internal error in svt_index_leaves():\n"
	      "    %s", leaf_errors[code]);
}

/* Recursive. 'span' = number of leaves under a node at depth 'ndim'. */
static void REC_index(SEXP SVT, const int *dim, int ndim, int64_t base,
		      SEXPTYPE Rtype, svt_leaf_index *ix)
{
	if (SVT == R_NilValue)
		return;
	if (ndim == 1) {
		int lac = 0;
		int code = index_leaf(SVT, Rtype, base, ix, &lac);
		if (code != 0)
			leaf_error(code);
		if (lac) ix->n_lacunar++;
		else     ix->n_regular++;
		return;
	}
	int SVT_len = dim[ndim - 1];
	if (!isVectorList(SVT) || LENGTH(SVT) != SVT_len)
		error("SparseArray internal error in svt_index_leaves():\n"
		      "    invalid SVT node");
	int64_t span = 1;
	for (int along = 1; along < ndim - 1; along++)
		span *= dim[along];
	for (int i = 0; i < SVT_len; i++)
		REC_index(VECTOR_ELT(SVT, i), dim, ndim - 1, base + i * span,
			  Rtype, ix);
}

void svt_index_leaves(SEXP SVT, const int *dim, int ndim, SEXPTYPE Rtype,
		      svt_leaf_index *ix)
{
	int64_t nleaf = 1;
	for (int along = 1; along < ndim; along++)
		nleaf *= dim[along];
	ix->nrow = dim[0];
	ix->nleaf = nleaf;
	ix->n_regular = ix->n_lacunar = 0;
	ix->leaf_ptr = (int64_t *) R_alloc(nleaf + 1, sizeof(int64_t));
	ix->offs = (const int **) R_alloc(nleaf > 0 ? nleaf : 1,
					  sizeof(const int *));
	ix->vals = (const void **) R_alloc(nleaf > 0 ? nleaf : 1,
					   sizeof(const void *));
	memset(ix->leaf_ptr, 0, sizeof(int64_t) * (size_t) (nleaf + 1));
	memset(ix->offs, 0, sizeof(const int *) * (size_t) nleaf);
	memset(ix->vals, 0, sizeof(const void *) * (size_t) nleaf);
	if (nleaf > 0 && ndim == 2 && SVT != R_NilValue) {
		/* a matrix: the columns in parallel */
		if (!isVectorList(SVT) || LENGTH(SVT) != dim[1])
			error("SparseArray internal error in "
			      "svt_index_leaves():\n    invalid SVT node");
		int bad = 0;
		int64_t n_lac = 0, n_reg = 0;
		#pragma omp parallel for schedule(static) \
			reduction(max:bad) reduction(+:n_lac, n_reg)
		for (int64_t j = 0; j < nleaf; j++) {
			SEXP leaf = VECTOR_ELT(SVT, j);
			if (leaf == R_NilValue)
				continue;
			int lac = 0;
			int code = index_leaf(leaf, Rtype, j, ix, &lac);
			if (code > bad)
				bad = code;
			if (code == 0) {
				if (lac) n_lac++;
				else     n_reg++;
			}
		}
		if (bad != 0)
			leaf_error(bad);
		ix->n_lacunar = n_lac;
		ix->n_regular = n_reg;
	} else if (nleaf > 0) {
		REC_index(SVT, dim, ndim, 0, Rtype, ix);
	}
	/* counts -> offsets */
	for (int64_t l = 0; l < nleaf; l++)
		ix->leaf_ptr[l + 1] += ix->leaf_ptr[l];
	ix->nnz = ix->leaf_ptr[nleaf];
}

static double now_ms(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* first leaf whose range ends after nonzero 'e' */
static int64_t leaf_holding(const int64_t *leaf_ptr, int64_t nleaf, int64_t e)
{
	int64_t lo = 0, hi = nleaf;
	while (lo < hi) {
		int64_t mid = lo + ((hi - lo) >> 1);
		if (leaf_ptr[mid + 1] <= e) lo = mid + 1;
		else                        hi = mid;
	}
	return lo;
}

/* ---- narrowing copies (fewer bytes over PCIe; widened again in HBM) ---- */

/* runtime-dispatched clones: R builds packages for the baseline ISA */
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define SVT_CLONES __attribute__((target_clones("arch=x86-64-v4", "avx2", "default")))
#else
#define SVT_CLONES
#endif

/* Row offsets are validated while they are copied: the kernels index
   shared-memory cells and result rows with them, so an offset outside
   [0, nrow) -- or a leaf whose offsets do not ascend strictly (two entries of
   one row) -- must never reach the device.  (The reference trusts its leaves;
   a malformed one misindexes host memory there.)  Returns nonzero when bad. */
SVT_CLONES
static int narrow_offs16(uint16_t *dst, const int *src, size_t n, int nrow,
			 int prev)
{
	unsigned bad = n > 0 && src[0] <= prev;
	for (size_t k = 0; k < n; k++) {
		const int o = src[k];
		bad |= (unsigned) o >= (unsigned) nrow;
		dst[k] = (uint16_t) o;
	}
	for (size_t k = 1; k < n; k++)      /* (separate: both loops vectorise) */
		bad |= src[k] <= src[k - 1];
	return bad != 0;
}

/* low halves only, for matrices with more than 65536 rows
   (SVTGPU_OFFS_U16_WRAPPED): bit 0 = bad offsets as above, bit 1 = some step
   inside the leaf (or the leaf's first offset) does not fit 16 bits */
SVT_CLONES
static int narrow_offs16w(uint16_t *dst, const int *src, size_t n, int nrow,
			  int prev)
{
	unsigned bad = n > 0 && src[0] <= prev;
	unsigned unfit = n > 0 && (prev < 0 ? src[0] > 65535
					    : src[0] - prev > 65535);
	for (size_t k = 0; k < n; k++) {
		const int o = src[k];
		bad |= (unsigned) o >= (unsigned) nrow;
		dst[k] = (uint16_t) o;
	}
	for (size_t k = 1; k < n; k++) {
		bad |= src[k] <= src[k - 1];
		unfit |= (unsigned) (src[k] - src[k - 1]) > 65535u;
	}
	return (bad != 0) | ((unfit != 0) << 1);
}

SVT_CLONES
static int copy_offs32(int32_t *dst, const int *src, size_t n, int nrow,
		       int prev)
{
	unsigned bad = n > 0 && src[0] <= prev;
	for (size_t k = 0; k < n; k++) {
		const int o = src[k];
		bad |= (unsigned) o >= (unsigned) nrow;
		dst[k] = o;
	}
	for (size_t k = 1; k < n; k++)      /* (separate: both loops vectorise) */
		bad |= src[k] <= src[k - 1];
	return bad != 0;
}

/* int32 -> int8 with NA -> -128; returns nonzero if some value is outside
   [-127, 127] */
SVT_CLONES
static int narrow_int8(int8_t *dst, const int *src, size_t n)
{
	unsigned bad = 0;
	for (size_t k = 0; k < n; k++) {
		const int v = src[k];
		const int is_na = v == NA_INTEGER;
		/* (unsigned arithmetic: v + 127 overflows for v near INT_MAX) */
		bad |= (((unsigned) v + 127u) > 254u) & !is_na;
		dst[k] = (int8_t) (is_na ? -128 : v);
	}
	return bad != 0;
}

/* double -> int8 when the value is an integer in [-127, 127]; NA_real_ ->
   -128; anything else (fractions, NaN, Inf, big values) reports failure */
SVT_CLONES
static int narrow_dbl8(int8_t *dst, const double *src, size_t n)
{
	unsigned bad = 0;
	for (size_t k = 0; k < n; k++) {
		const double v = src[k];
		if (v >= -127.0 && v <= 127.0) {
			const int8_t q = (int8_t) v;
			bad |= (double) q != v;
			dst[k] = q;
		} else if (R_IsNA(v)) {
			dst[k] = -128;
		} else {
			bad = 1;
			dst[k] = 0;
		}
	}
	return bad != 0;
}

static int narrowing_enabled(void)
{
	const char *v = getenv("SVTGPU_NARROW");
	return !(v != NULL && v[0] == '0');
}

int svt_upload_leaves(const svt_leaf_index *ix, SEXPTYPE Rtype, int want_offs,
		      int want_vals, svtgpu_matrix **out, double *flatten_ms)
{
	const int has_vals = want_vals && ix->n_regular > 0;
	const int flags = (want_offs ? SVTGPU_HAS_OFFS : 0) |
			  (has_vals ? SVTGPU_HAS_VALS : 0);
	const size_t vsz = Rtype == REALSXP ? sizeof(double) : sizeof(int);
	svtgpu_matrix *m = NULL;
	*out = NULL;
	*flatten_ms = 0.0;
	int rc = svtgpu_matrix_create(&m, ix->nrow, ix->nleaf, ix->nnz,
				      (int) Rtype, flags);
	if (rc != SVTGPU_OK)
		return rc;
	rc = svtgpu_matrix_set_leaf_ptr(m, ix->leaf_ptr);
	int64_t cap = 0;
	if (rc == SVTGPU_OK && ix->nnz > 0 && flags != 0)
		rc = svtgpu_matrix_stage_capacity(m, &cap);
	/* a matrix that fits one slot still goes in >= 4 pieces, so that the
	   copy engine works on one piece while the next is being flattened */
	if (rc == SVTGPU_OK && cap > 0) {
		int64_t piece = (ix->nnz + 3) / 4;
		if (piece < (1 << 18))
			piece = 1 << 18;
		if (piece < cap)
			cap = piece;
	}
	/* offsets fit 16 bits when nrow <= 65536; values are tried as int8 until
	   a slot holds one that does not fit */
	const int narrow = narrowing_enabled();
	const int offs16 = narrow && ix->nrow <= 65536;
	/* more rows: the low halves of the offsets, as long as every step
	   inside a leaf fits 16 bits (SVTGPU_OFFS_U16_WRAPPED) */
	int try_offs16w = narrow && !offs16;
	int try_vals8 = narrow;
	for (int64_t e0 = 0; rc == SVTGPU_OK && flags != 0 && e0 < ix->nnz;
	     e0 += cap) {
		const int64_t e1 = ix->nnz - e0 < cap ? ix->nnz : e0 + cap;
		int32_t *so = NULL;
		void *sv = NULL;
		rc = svtgpu_matrix_stage(m, e1 - e0, &so, &sv);
		if (rc != SVTGPU_OK)
			break;
		const int64_t l_first = leaf_holding(ix->leaf_ptr, ix->nleaf,
						     e0);
		const int64_t l_last = leaf_holding(ix->leaf_ptr, ix->nleaf,
						    e1 - 1);
		const double t0 = now_ms();
		int bad_offs = 0;
		int vals8 = try_vals8 && sv != NULL;
		int offs16w = try_offs16w && so != NULL;
		int redo_offs = 0, redo_vals = 0;
		for (int pass = 0; pass < 2; pass++) {
			/* pass 1 only redoes, in native width, what did not
			   fit its narrow form in pass 0 */
			if (pass == 1 && !redo_offs && !redo_vals)
				break;
			int bad = 0;
			#pragma omp parallel for schedule(dynamic, 64) \
				reduction(|:bad, bad_offs)
			for (int64_t l = l_first; l <= l_last; l++) {
				int64_t a = ix->leaf_ptr[l];
				int64_t b = ix->leaf_ptr[l + 1];
				if (a == b)
					continue;
				const int64_t from = a < e0 ? e0 : a;
				const int64_t to = b > e1 ? e1 : b;
				const size_t n = (size_t) (to - from);
				const size_t at = (size_t) (from - e0);
				if (so != NULL && (pass == 0 || redo_offs)) {
					const int *src = ix->offs[l] + (from - a);
					/* the entry before this piece of the leaf */
					const int prev = from > a ? src[-1] : -1;
					if (offs16)
						bad_offs |= narrow_offs16(
							(uint16_t *) so + at, src, n,
							(int) ix->nrow, prev);
					else if (offs16w && pass == 0)
						bad_offs |= narrow_offs16w(
							(uint16_t *) so + at, src, n,
							(int) ix->nrow, prev);
					else
						bad_offs |= copy_offs32(so + at, src,
							n, (int) ix->nrow, prev);
				}
				if (sv == NULL || (pass == 1 && !redo_vals))
					continue;
				const char *vsrc = ix->vals[l] == NULL ? NULL
					: (const char *) ix->vals[l] +
					  vsz * (size_t) (from - a);
				if (vals8 && pass == 0) {
					int8_t *dst = (int8_t *) sv + at;
					if (vsrc == NULL)
						memset(dst, 1, n);
					else if (Rtype == REALSXP)
						bad |= narrow_dbl8(dst,
							(const double *) vsrc, n);
					else
						bad |= narrow_int8(dst,
							(const int *) vsrc, n);
					continue;
				}
				char *dst = (char *) sv + vsz * at;
				if (vsrc != NULL) {
					memcpy(dst, vsrc, vsz * n);
				} else if (Rtype == REALSXP) {
					for (size_t k = 0; k < n; k++)
						((double *) dst)[k] = 1.0;
				} else {
					for (size_t k = 0; k < n; k++)
						((int *) dst)[k] = 1;
				}
			}
			if (pass == 0) {
				if (vals8 && bad) {
					vals8 = 0;  /* redo this slot's values ... */
					redo_vals = 1;
				}
				if (offs16w && (bad_offs & 2)) {
					offs16w = 0;    /* ... or its offsets ... */
					redo_offs = 1;
				}
				bad_offs &= 1;
				continue;
			}
			break;
		}
		if (try_vals8 && sv != NULL && !vals8)
			try_vals8 = 0;          /* ... and stop trying */
		if (try_offs16w && so != NULL && !offs16w)
			try_offs16w = 0;
		*flatten_ms += now_ms() - t0;
		if (bad_offs) {
			svtgpu_matrix_free(m);
			return SVT_FLATTEN_BAD_OFFSETS;
		}
		rc = svtgpu_matrix_commit_packed(m, e0, e1 - e0,
				offs16 ? 2 : offs16w ? SVTGPU_OFFS_U16_WRAPPED : 4,
				vals8 ? 1 : (int) vsz);
	}
	if (rc == SVTGPU_OK)
		rc = svtgpu_matrix_finish_upload(m);
	if (rc != SVTGPU_OK) {
		svtgpu_matrix_free(m);
		return rc;
	}
	*out = m;
	return SVTGPU_OK;
}
