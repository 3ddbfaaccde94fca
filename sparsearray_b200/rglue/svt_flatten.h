/* Flattening of an SVT (R nested list of leaves) into the device CSC of
 * include/svtgpu.h.  Host side of north-star step (1); the reference's nearest
 * relative is dump_SVT_to_CsparseMatrix_slots(),
 * src/SVT_SparseArray_class.c:598-679. */
#ifndef SVT_FLATTEN_H
#define SVT_FLATTEN_H

#include <Rdefines.h>
#include <stdint.h>

#include "../../include/svtgpu.h"

typedef struct svt_leaf_index {
	int64_t nrow;           /* dim[0] */
	int64_t nleaf;          /* prod(dim[-1]) */
	int64_t nnz;
	int64_t *leaf_ptr;      /* nleaf + 1 (R_alloc) */
	const int **offs;       /* per leaf: nzoffs payload, NULL if empty */
	const void **vals;      /* per leaf: nzvals payload, NULL if empty/lacunar */
	int64_t n_regular;      /* leaves carrying nzvals */
	int64_t n_lacunar;      /* non-empty leaves with nzvals == NULL */
} svt_leaf_index;

/* Walk the tree, validate every leaf the way unzip_leaf()/toSparseVec() do
 * (src/leaf_utils.h:88-132, src/SparseVec.h:44-88) and record payload
 * pointers.  All memory comes from R_alloc(); raises an R error on an invalid
 * SVT.  Must run on the R main thread. */
void svt_index_leaves(SEXP SVT, const int *dim, int ndim, SEXPTYPE Rtype,
		      svt_leaf_index *ix);

/* Create the device matrix and stream the leaves into it through the pinned
 * staging slots (OpenMP threads copy leaf payloads; no R API inside).
 * want_offs/want_vals select which arrays the coming operation needs; an SVT
 * whose non-empty leaves are all lacunar gets no value array at all, a mixed
 * SVT has ones materialised for its lacunar leaves.
 * Returns an svtgpu status; *flatten_ms = host time spent copying leaves. */
/* returned by svt_upload_leaves() when a leaf holds a row offset outside
 * [0, nrow) or offsets that do not ascend strictly */
#define SVT_FLATTEN_BAD_OFFSETS (-7)

int svt_upload_leaves(const svt_leaf_index *ix, SEXPTYPE Rtype, int want_offs,
		      int want_vals, svtgpu_matrix **out, double *flatten_ms);

#endif  /* SVT_FLATTEN_H */
