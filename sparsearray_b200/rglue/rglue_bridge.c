/* CSC <-> device-resident SVT bridges.
 *
 * TENxMatrix / HDF5 loaders and Matrix::dgCMatrix objects hold a sparse
 * matrix as CSC arrays (indptr / indices / data); the reference turns those
 * into an SVT with C_build_SVT_from_CSC() and back with
 * C_from_SVT_SparseMatrix_to_CsparseMatrix()
 * (src/SVT_SparseArray_class.c:751-861 and :598-679) -- one R vector pair per
 * column.  The device CSC of include/svtgpu.h IS that layout, so a matrix
 * that is only going to be summarised / multiplied never needs the list:
 *
 *   C_svtgpu_from_CSC(dim, indptr, data, indices, indices_are_1based)
 *       -> device-resident handle (usable wherever x@SVT is), with the
 *          reference's conventions: explicit zeros in 'data' are dropped
 *          (build_leaf_from_CsparseMatrix_col() :715-742), the entries of a
 *          column are ordered by row (_INPLACE_order_leaf_by_nzoff() via
 *          :797-800), 'indptr' may be integer or double (:754-755).  Unlike
 *          the reference, row indices are validated (range, duplicates): the
 *          kernels index shared-memory cells with them.
 *   C_svtgpu_to_CSC(handle, as_ngCMatrix)
 *       -> list(p, i, x) exactly as C_from_SVT_SparseMatrix_to_CsparseMatrix()
 *          returns it, downloaded from HBM; refuses more than INT_MAX
 *          nonzeros with the reference's message (:643-647).
 *
 * These are extension entry points with their own names: the package's own
 * C_build_SVT_from_CSC / C_from_SVT_... stay what they are.
 */
#include "rglue_common.h"

#include <limits.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#undef match
#include <omp.h>
#endif

#define SLOTP(slotp, j) \
	(IS_INTEGER(slotp) ? (int64_t) INTEGER(slotp)[j] : (int64_t) REAL(slotp)[j])

static inline int is_nonzero(SEXPTYPE t, const void *data, int64_t k)
{
	if (t == REALSXP)
		return ((const double *) data)[k] != 0.0;   /* NaN / NA: nonzero */
	return ((const int *) data)[k] != 0;
}

/* order the (off, val) pairs of one column by off (insertion sort: columns
   that arrive unsorted are rare and short) */
static void sort_column(int32_t *offs, void *vals, size_t vsz, int64_t n)
{
	for (int64_t a = 1; a < n; a++) {
		const int32_t o = offs[a];
		unsigned char v[8];
		memcpy(v, (char *) vals + vsz * (size_t) a, vsz);
		int64_t b = a - 1;
		while (b >= 0 && offs[b] > o) {
			offs[b + 1] = offs[b];
			memcpy((char *) vals + vsz * (size_t) (b + 1),
			       (char *) vals + vsz * (size_t) b, vsz);
			b--;
		}
		offs[b + 1] = o;
		memcpy((char *) vals + vsz * (size_t) (b + 1), v, vsz);
	}
}

SEXP C_svtgpu_from_CSC(SEXP dim, SEXP indptr, SEXP data, SEXP indices,
		       SEXP indices_are_1based)
{
	if (!(IS_INTEGER(dim) && LENGTH(dim) == 2))
		error("SparseArray internal error in C_svtgpu_from_CSC():\n"
		      "    invalid 'dim'");
	const int nrow = INTEGER(dim)[0];
	const int ncol = INTEGER(dim)[1];
	if (nrow < 0 || ncol < 0)
		error("SparseArray internal error in C_svtgpu_from_CSC():\n"
		      "    invalid 'dim'");
	const SEXPTYPE Rtype = TYPEOF(data);
	if (Rtype != LGLSXP && Rtype != INTSXP && Rtype != REALSXP)
		error("SparseArray objects of type() \"%s\" are not "
		      "supported by the SparseArray GPU path",
		      type2char(Rtype));
	if (!(IS_INTEGER(indices) && XLENGTH(indices) == XLENGTH(data)))
		error("SparseArray internal error in C_svtgpu_from_CSC():\n"
		      "    invalid 'indices'");
	if (!(IS_LOGICAL(indices_are_1based) &&
	      LENGTH(indices_are_1based) == 1))
		error("'indices_are_1based' must be TRUE or FALSE");
	const int one_based = LOGICAL(indices_are_1based)[0] != 0;
	if (!((IS_INTEGER(indptr) || IS_NUMERIC(indptr)) &&
	      LENGTH(indptr) == ncol + 1 && SLOTP(indptr, 0) == 0))
		error("SparseArray internal error in build_SVT_from_CSC():\n"
		      "    invalid 'slotp'");
	const int64_t ix_len = SLOTP(indptr, ncol);
	if (ix_len != (int64_t) XLENGTH(indices))
		error("SparseArray internal error in build_SVT_from_CSC():\n"
		      "    invalid 'slotp'");
	const int *idx = INTEGER(indices);
	const void *dat = DATAPTR(data);
	const size_t vsz = Rtype == REALSXP ? sizeof(double) : sizeof(int);

	/* pass 1: nonzeros per column (zeros in 'data' are dropped) */
	int64_t *leaf_ptr = (int64_t *) R_alloc((size_t) ncol + 1,
						sizeof(int64_t));
	leaf_ptr[0] = 0;
	int bad = 0;
	#pragma omp parallel for schedule(static) reduction(|:bad)
	for (int j = 0; j < ncol; j++) {
		const int64_t a = SLOTP(indptr, j), b = SLOTP(indptr, j + 1);
		int64_t n = 0;
		if (a < 0 || b < a || b > ix_len) {
			bad |= 1;
		} else {
			for (int64_t k = a; k < b; k++)
				n += is_nonzero(Rtype, dat, k);
		}
		leaf_ptr[j + 1] = n;
	}
	if (bad)
		error("SparseArray internal error in build_SVT_from_CSC():\n"
		      "    invalid 'slotp'");
	for (int j = 0; j < ncol; j++) {
		if (leaf_ptr[j + 1] > nrow)
			bad = 1;
		leaf_ptr[j + 1] += leaf_ptr[j];
	}
	const int64_t nnz = leaf_ptr[ncol];
	int32_t *offs = NULL;
	void *vals = NULL;
	if (nnz > 0) {
		offs = (int32_t *) malloc(sizeof(int32_t) * (size_t) nnz);
		vals = malloc(vsz * (size_t) nnz);
		if (offs == NULL || vals == NULL) {
			free(offs);
			free(vals);
			error("C_svtgpu_from_CSC(): out of memory");
		}
	}
	/* pass 2: compact, make the row indices 0-based, validate, order */
	#pragma omp parallel for schedule(dynamic, 256) reduction(|:bad)
	for (int j = 0; j < ncol; j++) {
		const int64_t a = SLOTP(indptr, j), b = SLOTP(indptr, j + 1);
		int64_t w = leaf_ptr[j];
		const int64_t w0 = w;
		int sorted = 1;
		for (int64_t k = a; k < b; k++) {
			if (!is_nonzero(Rtype, dat, k))
				continue;
			const int64_t o = (int64_t) idx[k] - one_based;
			if (idx[k] == NA_INTEGER || o < 0 || o >= nrow) {
				bad |= 2;
				continue;
			}
			if (w > w0 && (int32_t) o <= offs[w - 1])
				sorted = 0;
			offs[w] = (int32_t) o;
			memcpy((char *) vals + vsz * (size_t) w,
			       (const char *) dat + vsz * (size_t) k, vsz);
			w++;
		}
		if (bad & 2)
			continue;
		if (!sorted) {
			sort_column(offs + w0, (char *) vals + vsz * (size_t) w0,
				    vsz, w - w0);
			for (int64_t k = w0 + 1; k < w; k++)
				if (offs[k] == offs[k - 1])
					bad |= 4;
		}
	}
	if (bad) {
		free(offs);
		free(vals);
		if (bad & 2)
			error("C_svtgpu_from_CSC(): 'indices' contains row "
			      "indices outside the matrix");
		if (bad & 4)
			error("C_svtgpu_from_CSC(): 'indices' contains "
			      "duplicates within a column");
		error("C_svtgpu_from_CSC(): a column holds more entries than "
		      "the matrix has rows");
	}
	svtgpu_matrix *m = NULL;
	int rc = svtgpu_matrix_create(&m, nrow, ncol, nnz, (int) Rtype,
			nnz > 0 ? (SVTGPU_HAS_OFFS | SVTGPU_HAS_VALS) : 0);
	if (rc == SVTGPU_OK)
		rc = svtgpu_matrix_upload(m, leaf_ptr, offs, vals);
	if (rc == SVTGPU_OK)
		rc = svtgpu_matrix_finish_upload(m);
	free(offs);
	free(vals);
	if (rc != SVTGPU_OK) {
		if (m != NULL)
			svtgpu_matrix_free(m);
		rglue_fail(rc, "svtgpu_matrix_upload");
	}
	rglue_record_timings(m, 0.0);
	return rglue_make_handle(m);
}

SEXP C_svtgpu_to_CSC(SEXP handle, SEXP as_ngCMatrix)
{
	if (TYPEOF(handle) != EXTPTRSXP)
		error("'handle' must be a device-resident SVT handle");
	svtgpu_matrix *m = (svtgpu_matrix *) R_ExternalPtrAddr(handle);
	if (m == NULL)
		error("the device-resident SVT handle has been released");
	if (!(IS_LOGICAL(as_ngCMatrix) && LENGTH(as_ngCMatrix) == 1))
		error("'as.ngCMatrix' must be TRUE or FALSE");
	int64_t nrow = 0, nleaf = 0, nnz = 0;
	int val_type = 0, flags = 0;
	if (svtgpu_matrix_info(m, &nrow, &nleaf, &nnz, &val_type, &flags) !=
	    SVTGPU_OK)
		rglue_fail(SVTGPU_ERR_ARG, "svtgpu_matrix_info");
	if (nleaf > INT_MAX)
		error("object to coerce to [d|l]gCMatrix "
		      "must have exactly 2 dimensions");
	if (nnz > INT_MAX)
		error("SVT_SparseMatrix object contains too many nonzero "
		      "values to be turned into a dgCMatrix or lgCMatrix "
		      "object");
	if (nnz > 0 && !(flags & SVTGPU_HAS_OFFS))
		error("the device-resident SVT handle holds no row offsets");
	const int drop_x = LOGICAL(as_ngCMatrix)[0];
	SEXP ans = PROTECT(NEW_LIST(3));
	SEXP slotp = SET_VECTOR_ELT(ans, 0, NEW_INTEGER(nleaf + 1));
	SEXP sloti = SET_VECTOR_ELT(ans, 1, NEW_INTEGER(nnz));
	SEXP slotx = R_NilValue;
	if (!drop_x)
		slotx = SET_VECTOR_ELT(ans, 2,
				       allocVector((SEXPTYPE) val_type, nnz));
	int64_t *lp = (int64_t *) R_alloc((size_t) nleaf + 1, sizeof(int64_t));
	const int has_vals = (flags & SVTGPU_HAS_VALS) != 0;
	int rc = svtgpu_matrix_download(m, lp, nnz > 0 ? INTEGER(sloti) : NULL,
			(!drop_x && has_vals && nnz > 0) ? DATAPTR(slotx) : NULL);
	if (rc != SVTGPU_OK) {
		UNPROTECT(1);
		rglue_fail(rc, "svtgpu_matrix_download");
	}
	for (int64_t j = 0; j <= nleaf; j++)
		INTEGER(slotp)[j] = (int) lp[j];
	if (!drop_x && !has_vals) {   /* every leaf lacunar: the values are ones */
		if (val_type == REALSXP)
			for (int64_t k = 0; k < nnz; k++)
				REAL(slotx)[k] = 1.0;
		else
			for (int64_t k = 0; k < nnz; k++)
				INTEGER(slotx)[k] = 1;
	}
	rglue_record_timings(m, 0.0);
	UNPROTECT(1);
	return ans;
}
