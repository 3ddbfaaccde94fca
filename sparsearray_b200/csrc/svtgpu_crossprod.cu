/* SVT x dense products on a device CSC.
 *
 *   crossprod(svt, Y):  ans[l, k] = sum_i svt[i, l] * Y[i, k]   (gather)
 *   svt %*% D:          ans[i, k] = sum_l svt[i, l] * D[l, k]   (scatter)
 *
 * Replaces crossprod2_SVT_mat_{double,int}() / crossprod2_mat_SVT_*()
 * (src/SparseMatrix_mult.c:385-547), which walk the whole SVT once per dense
 * column, and -- for `%*%` -- the transpose the R method performs first
 * (R/SparseMatrix-mult.R:196-198, src/SparseArray_aperm.c:348-423).  Here the
 * SVT is streamed from HBM once for all K dense columns.
 *
 * The dense operand is kept row-major (K contiguous doubles per row) so the
 * K values a nonzero needs are one coalesced segment; it is small enough
 * (nrow x K x 8 B) to live in L2.  The kernels are bound by that on-chip
 * gather (K x 8 B per nonzero against 12 B from HBM), not by HBM and not by
 * FP64 issue; there is no dense contraction here, so no tensor cores.
 *
 * NA rules follow the reference per dense column (svt_dot_finalize()).
 */
#include "svtgpu_internal.h"
#include "svt_ptx.cuh"

#include <string.h>

namespace {

/* dense operand: column-major int32/double (optionally transposed) ->
 * row-major double [n x K]; one thread per output element */
template <typename T>
__global__ void __launch_bounds__(256)
dense_to_rowmajor(const T *__restrict__ y, int64_t y_nrow, int64_t n,
		  int64_t K, int transpose, double *__restrict__ out)
{
	const int64_t total = n * K;
	const int64_t stride = (int64_t) gridDim.x * blockDim.x;
	for (int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	     t < total; t += stride) {
		const int64_t i = t / K, k = t - i * K;
		const T v = transpose ? y[k + i * y_nrow] : y[i + k * y_nrow];
		out[t] = (double) v;
	}
}

/* per dense column: how many entries are non-finite / NA */
template <typename T>
__global__ void __launch_bounds__(256)
dense_col_info(const T *__restrict__ y, int64_t y_nrow, int64_t n, int64_t K,
	       int transpose, SvtDenseColInfo *info)
{
	const int64_t k = blockIdx.x;
	int nf = 0, na = 0;
	for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
		const T v = transpose ? y[k + i * y_nrow] : y[i + k * y_nrow];
		if (sizeof(T) == 4) {
			const int bad = (int32_t) v == SVT_NA_INT;
			nf += bad; na += bad;
		} else {
			const double d = (double) v;
			nf += !svt_isfinite(d);
			na += svt_is_na_real(d);
		}
	}
	nf = (int) svt_warp_sum((long long) nf);
	na = (int) svt_warp_sum((long long) na);
	if ((threadIdx.x & 31) == 0) {
		if (nf) atomicAdd(&info[k].n_nonfinite, nf);
		if (na) atomicAdd(&info[k].n_na, na);
	}
}

__device__ __forceinline__ bool val_is_na(int32_t x) { return x == SVT_NA_INT; }
__device__ __forceinline__ bool val_is_na(double x) { return svt_is_na_real(x); }

#define CP_WARPS 8

/* Gather: one warp per leaf, lanes own dense columns k0+lane and k0+32+lane.
 * SVT_LEFT: ans is nleaf x K column-major (staged through shared memory so
 * the 8 leaves of a CTA are written as contiguous runs); else K x nleaf. */
template <typename T, bool LACUNAR, bool CHECK_NF, bool SVT_LEFT>
__global__ void __launch_bounds__(CP_WARPS * 32)
crossprod_gather(const int32_t *__restrict__ offs, const T *__restrict__ vals,
		 const int64_t *__restrict__ leaf_ptr, int64_t nleaf,
		 const double *__restrict__ Y, int64_t K, int is_double,
		 const SvtDenseColInfo *__restrict__ info,
		 double *__restrict__ ans)
{
	__shared__ double tile[64][CP_WARPS + 1];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int64_t ngroups = (nleaf + CP_WARPS - 1) / CP_WARPS;
	for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
		const int64_t leaf = g * CP_WARPS + warp;
		int64_t start = 0, end = 0;
		if (leaf < nleaf) {
			start = leaf_ptr[leaf];
			end = leaf_ptr[leaf + 1];
		}
		for (int64_t k0 = 0; k0 < K; k0 += 64) {
			const int64_t ka = k0 + lane, kb = k0 + 32 + lane;
			const bool ha = ka < K, hb = kb < K;
			double sa = 0.0, sb = 0.0;
			int hits_a = 0, hits_b = 0, leaf_na = 0;
			for (int64_t e0 = start; e0 < end; e0 += 32) {
				const int64_t e = e0 + lane;
				int32_t my_off = 0;
				double my_v = 0.0;
				if (e < end) {
					my_off = offs[e];
					if (LACUNAR) {
						my_v = 1.0;
					} else {
						const T x = vals[e];
						leaf_na |= val_is_na(x);
						my_v = (double) x;
					}
				}
				const int n = (int) (end - e0 < 32 ? end - e0
								   : 32);
#pragma unroll 8
				for (int i = 0; i < n; i++) {
					const int64_t off = __shfl_sync(
						SVT_FULL_MASK, my_off, i);
					const double v = __shfl_sync(
						SVT_FULL_MASK, my_v, i);
					const double *yr = Y + off * K;
					if (ha) {
						const double y = yr[ka];
						sa += v * y;
						if (CHECK_NF)
							hits_a += !svt_isfinite(y);
					}
					if (hb) {
						const double y = yr[kb];
						sb += v * y;
						if (CHECK_NF)
							hits_b += !svt_isfinite(y);
					}
				}
			}
			leaf_na = __any_sync(SVT_FULL_MASK, leaf_na);
			if (ha)
				sa = svt_dot_finalize(is_double, sa, leaf_na,
						      hits_a, info[ka]);
			if (hb)
				sb = svt_dot_finalize(is_double, sb, leaf_na,
						      hits_b, info[kb]);
			if (!SVT_LEFT) {
				if (leaf < nleaf) {
					if (ha) ans[ka + leaf * K] = sa;
					if (hb) ans[kb + leaf * K] = sb;
				}
				continue;
			}
			tile[lane][warp] = sa;
			tile[32 + lane][warp] = sb;
			__syncthreads();
			/* 8 consecutive leaves of one dense column = 64 B */
			for (int t = threadIdx.x; t < 64 * CP_WARPS;
			     t += CP_WARPS * 32) {
				const int kk = t / CP_WARPS, w = t % CP_WARPS;
				const int64_t l = g * CP_WARPS + w;
				if (k0 + kk < K && l < nleaf)
					ans[l + (k0 + kk) * nleaf] = tile[kk][w];
			}
			__syncthreads();
		}
	}
}

/* Scatter (svt %*% D): one warp per leaf l, lanes own columns k; the leaf's
 * dense row D[l, ] stays in registers and every nonzero adds v * D[l, k] into
 * the row-major product with fp64 reductions in L2. */
template <typename T, bool LACUNAR, bool CHECK_NF>
__global__ void __launch_bounds__(256)
matmul_scatter(const int32_t *__restrict__ offs, const T *__restrict__ vals,
	       const int64_t *__restrict__ leaf_ptr, int64_t nleaf,
	       const double *__restrict__ D, int64_t K,
	       double *__restrict__ prod,       /* nrow x K row-major */
	       int32_t *__restrict__ row_na,    /* nrow: row holds an NA */
	       int32_t *__restrict__ hits)      /* nrow x K, CHECK_NF only */
{
	const int lane = threadIdx.x & 31;
	const int64_t warps = ((int64_t) gridDim.x * blockDim.x) >> 5;
	const int64_t gw = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	for (int64_t leaf = gw; leaf < nleaf; leaf += warps) {
		const int64_t start = leaf_ptr[leaf], end = leaf_ptr[leaf + 1];
		if (start == end)
			continue;
		for (int64_t k0 = 0; k0 < K; k0 += 64) {
			const int64_t ka = k0 + lane, kb = k0 + 32 + lane;
			const bool ha = ka < K, hb = kb < K;
			const double da = ha ? D[leaf * K + ka] : 0.0;
			const double db = hb ? D[leaf * K + kb] : 0.0;
			const bool nfa = CHECK_NF && ha && !svt_isfinite(da);
			const bool nfb = CHECK_NF && hb && !svt_isfinite(db);
			for (int64_t e0 = start; e0 < end; e0 += 32) {
				const int64_t e = e0 + lane;
				int32_t my_off = 0;
				double my_v = 0.0;
				if (e < end) {
					my_off = offs[e];
					if (LACUNAR) {
						my_v = 1.0;
					} else {
						const T x = vals[e];
						if (k0 == 0 && val_is_na(x))
							row_na[my_off] = 1;
						my_v = (double) x;
					}
				}
				const int n = (int) (end - e0 < 32 ? end - e0
								   : 32);
				for (int i = 0; i < n; i++) {
					const int64_t off = __shfl_sync(
						SVT_FULL_MASK, my_off, i);
					const double v = __shfl_sync(
						SVT_FULL_MASK, my_v, i);
					double *pr = prod + off * K;
					if (ha) atomicAdd(pr + ka, v * da);
					if (hb) atomicAdd(pr + kb, v * db);
					if (nfa) atomicAdd(hits + off * K + ka, 1);
					if (nfb) atomicAdd(hits + off * K + kb, 1);
				}
			}
		}
	}
}

/* row-major product + flags -> column-major answer with the NA rules */
__global__ void __launch_bounds__(256)
matmul_finalize(const double *__restrict__ prod,
		const int32_t *__restrict__ row_na,
		const int32_t *__restrict__ hits, int64_t nrow, int64_t K,
		int is_double, const SvtDenseColInfo *__restrict__ info,
		double *__restrict__ ans, int ans_rowmajor)
{
	const int64_t total = nrow * K;
	const int64_t stride = (int64_t) gridDim.x * blockDim.x;
	for (int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	     t < total; t += stride) {
		const int64_t i = t / K, k = t - i * K;
		const double v = svt_dot_finalize(is_double, prod[t],
				row_na[i], hits != NULL ? hits[t] : 0, info[k]);
		if (ans_rowmajor) ans[t] = v;
		else              ans[i + k * nrow] = v;
	}
}

inline unsigned grid_for(int64_t n, int per_block)
{
	int64_t b = (n + per_block - 1) / per_block;
	int64_t cap = (int64_t) svtgpu_sm_count() * 16;
	if (b > cap) b = cap;
	if (b < 1) b = 1;
	return (unsigned) b;
}

template <typename T, bool LAC>
int launch_gather(const svtgpu_matrix *m, const double *Y, int64_t K,
		  const SvtDenseColInfo *info, bool check_nf, bool svt_left,
		  double *ans, cudaStream_t s)
{
	const int64_t ngroups = (m->nleaf + CP_WARPS - 1) / CP_WARPS;
	unsigned grid = (unsigned) (ngroups < (int64_t) svtgpu_sm_count() * 8
			? (ngroups > 0 ? ngroups : 1)
			: (int64_t) svtgpu_sm_count() * 8);
	const int isd = svt_is_double(m->val_type);
#define GATHER(NF, LEFT) crossprod_gather<T, LAC, NF, LEFT> \
		<<<grid, CP_WARPS * 32, 0, s>>>(m->d_offs, \
		(const T *) m->d_vals, m->d_leaf_ptr, m->nleaf, Y, K, isd, \
		info, ans)
	if (check_nf) { if (svt_left) GATHER(true, true);
			else GATHER(true, false); }
	else          { if (svt_left) GATHER(false, true);
			else GATHER(false, false); }
#undef GATHER
	SVT_CUDA(cudaGetLastError());
	svtgpu_count_launch(1);
	return SVTGPU_OK;
}

template <typename T, bool LAC>
int launch_scatter(const svtgpu_matrix *m, const double *D, int64_t K,
		   bool check_nf, double *prod, int32_t *row_na,
		   int32_t *hits, cudaStream_t s)
{
	unsigned grid = grid_for(m->nleaf, 8);
	if (check_nf)
		matmul_scatter<T, LAC, true><<<grid, 256, 0, s>>>(m->d_offs,
			(const T *) m->d_vals, m->d_leaf_ptr, m->nleaf, D, K,
			prod, row_na, hits);
	else
		matmul_scatter<T, LAC, false><<<grid, 256, 0, s>>>(m->d_offs,
			(const T *) m->d_vals, m->d_leaf_ptr, m->nleaf, D, K,
			prod, row_na, hits);
	SVT_CUDA(cudaGetLastError());
	svtgpu_count_launch(1);
	return SVTGPU_OK;
}

/* Upload a column-major dense operand and lay it out row-major as doubles;
 * d_info receives the per-column NA / non-finite counts. */
int prepare_dense(const void *y, int y_type, int64_t y_nrow, int64_t y_ncol,
		  int transpose, int64_t n, int64_t K, double *d_rowmajor,
		  void *d_raw, SvtDenseColInfo *d_info, int *any_bad,
		  cudaStream_t s)
{
	const size_t esz = y_type == SVTGPU_DOUBLE ? 8 : 4;
	SVT_CUDA(cudaMemcpyAsync(d_raw, y, esz * (size_t) (y_nrow * y_ncol),
				 cudaMemcpyHostToDevice, s));
	SVT_CUDA(cudaMemsetAsync(d_info, 0, sizeof(SvtDenseColInfo) * K, s));
	if (y_type == SVTGPU_DOUBLE) {
		dense_to_rowmajor<double><<<grid_for(n * K, 256), 256, 0, s>>>(
			(const double *) d_raw, y_nrow, n, K, transpose,
			d_rowmajor);
		dense_col_info<double><<<(unsigned) K, 256, 0, s>>>(
			(const double *) d_raw, y_nrow, n, K, transpose,
			d_info);
	} else {
		dense_to_rowmajor<int32_t><<<grid_for(n * K, 256), 256, 0, s>>>(
			(const int32_t *) d_raw, y_nrow, n, K, transpose,
			d_rowmajor);
		dense_col_info<int32_t><<<(unsigned) K, 256, 0, s>>>(
			(const int32_t *) d_raw, y_nrow, n, K, transpose,
			d_info);
	}
	SVT_CUDA(cudaGetLastError());
	svtgpu_count_launch(2);
	/* the slow (non-finite) variant is only needed when some column of
	   the dense operand is not clean */
	SvtDenseColInfo *h = (SvtDenseColInfo *) malloc(
				sizeof(SvtDenseColInfo) * (size_t) K);
	if (h == NULL) {
		svtgpu_set_error("crossprod: out of host memory");
		return SVTGPU_ERR_NOMEM;
	}
	cudaError_t e = cudaMemcpyAsync(h, d_info, sizeof(SvtDenseColInfo) * K,
					cudaMemcpyDeviceToHost, s);
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(s);
	*any_bad = 0;
	for (int64_t k = 0; e == cudaSuccess && k < K; k++)
		if (h[k].n_nonfinite != 0)
			*any_bad = 1;
	free(h);
	SVT_CUDA(e);
	return SVTGPU_OK;
}

int check_product_types(const svtgpu_matrix *m, int dense_type,
			const char *what)
{
	SVT_ARG(m->val_type == SVTGPU_DOUBLE || m->val_type == SVTGPU_INT,
		"%s: input type is not supported yet (only \"double\" and "
		"\"integer\")", what);
	SVT_ARG(dense_type == m->val_type,
		"%s: the dense operand must have the type of the SVT", what);
	SVT_ARG((m->flags & SVTGPU_HAS_OFFS) || m->nnz == 0,
		"%s: the matrix was uploaded without row offsets", what);
	return SVTGPU_OK;
}

}  /* namespace */

extern "C" int svtgpu_crossprod_dev(svtgpu_matrix *m, const void *d_y_rowmajor,
				    int y_type, int64_t K, double *d_ans,
				    void *stream)
{
	SVT_CHECK(svtgpu_require_device());
	SVT_ARG(m != NULL && d_ans != NULL && (d_y_rowmajor != NULL ||
		m->nrow == 0 || K == 0), "svtgpu_crossprod_dev: NULL argument");
	SVT_ARG(y_type == SVTGPU_DOUBLE,
		"svtgpu_crossprod_dev: the device operand must be double");
	cudaStream_t s = (cudaStream_t) stream;
	if (m->nleaf == 0 || K == 0)
		return SVTGPU_OK;
	/* device form assumes a clean (finite) dense operand */
	void *scratch = NULL;
	SVT_CHECK(svtgpu_scratch(m, sizeof(SvtDenseColInfo) * (size_t) K,
				 &scratch));
	SVT_CUDA(cudaMemsetAsync(scratch, 0, sizeof(SvtDenseColInfo) * K, s));
	const SvtDenseColInfo *info = (const SvtDenseColInfo *) scratch;
	const double *Y = (const double *) d_y_rowmajor;
	if (!(m->flags & SVTGPU_HAS_VALS))
		return launch_gather<int32_t, true>(m, Y, K, info, false, true,
						    d_ans, s);
	if (svt_is_double(m->val_type))
		return launch_gather<double, false>(m, Y, K, info, false, true,
						    d_ans, s);
	return launch_gather<int32_t, false>(m, Y, K, info, false, true,
					     d_ans, s);
}

extern "C" int svtgpu_crossprod(svtgpu_matrix *m, const void *y, int y_type,
				int64_t y_nrow, int64_t y_ncol,
				int transpose_y, int svt_on_left, double *ans)
{
	SVT_CHECK(svtgpu_require_device());
	SVT_ARG(m != NULL && ans != NULL, "svtgpu_crossprod: NULL argument");
	SVT_CHECK(check_product_types(m, y_type, "crossprod"));
	const int64_t K = transpose_y ? y_nrow : y_ncol;
	const int64_t n = transpose_y ? y_ncol : y_nrow;
	SVT_ARG(n == m->nrow, "input objects are non-conformable");
	SVT_CHECK(svtgpu_matrix_finish_upload(m));
	m->tm.kernel_ms = m->tm.d2h_ms = 0.0;
	m->tm.d2h_bytes = 0.0;
	m->tm.launches = 0;
	const size_t nout = (size_t) (m->nleaf * K);
	if (nout == 0)
		return SVTGPU_OK;
	if (m->nnz == 0) {   /* x_SVT == NULL: src/SparseMatrix_mult.c:389-390 */
		memset(ans, 0, sizeof(double) * nout);
		return SVTGPU_OK;
	}
	cudaStream_t s = 0;
	const size_t esz = y_type == SVTGPU_DOUBLE ? 8 : 4;
	const size_t raw_bytes = (esz * (size_t) (y_nrow * y_ncol) + 255) &
				 ~(size_t) 255;
	const size_t rm_bytes = (8 * (size_t) (n * K) + 255) & ~(size_t) 255;
	const size_t info_bytes = (sizeof(SvtDenseColInfo) * (size_t) K + 255) &
				  ~(size_t) 255;
	char *d_buf = NULL;
	SVT_CUDA(cudaMallocAsync((void **) &d_buf, raw_bytes + rm_bytes +
				 info_bytes + 8 * nout, s));
	void *d_raw = d_buf;
	double *d_rm = (double *) (d_buf + raw_bytes);
	SvtDenseColInfo *d_info = (SvtDenseColInfo *) (d_buf + raw_bytes +
						       rm_bytes);
	double *d_ans = (double *) (d_buf + raw_bytes + rm_bytes + info_bytes);
	int any_bad = 0;
	int64_t l0 = svtgpu_launch_count();
	int rc = prepare_dense(y, y_type, y_nrow, y_ncol, transpose_y, n, K,
			       d_rm, d_raw, d_info, &any_bad, s);
	SvtTimer t;
	if (rc == SVTGPU_OK)
		rc = svt_timer_begin(&t, s);
	if (rc == SVTGPU_OK) {
		const bool left = svt_on_left != 0;
		if (!(m->flags & SVTGPU_HAS_VALS))
			rc = launch_gather<int32_t, true>(m, d_rm, K, d_info,
					any_bad, left, d_ans, s);
		else if (svt_is_double(m->val_type))
			rc = launch_gather<double, false>(m, d_rm, K, d_info,
					any_bad, left, d_ans, s);
		else
			rc = launch_gather<int32_t, false>(m, d_rm, K, d_info,
					any_bad, left, d_ans, s);
		int rc2 = svt_timer_end(&t, &m->tm.kernel_ms);
		if (rc == SVTGPU_OK)
			rc = rc2;
	}
	m->tm.launches = (int) (svtgpu_launch_count() - l0);
	if (rc == SVTGPU_OK) {
		rc = svt_timer_begin(&t, s);
		cudaError_t e = cudaMemcpyAsync(ans, d_ans, 8 * nout,
						cudaMemcpyDeviceToHost, s);
		int rc2 = svt_timer_end(&t, &m->tm.d2h_ms);
		if (e != cudaSuccess)
			rc = svtgpu_cuda_fail(e, "crossprod D2H", __FILE__,
					      __LINE__);
		else if (rc == SVTGPU_OK)
			rc = rc2;
		m->tm.d2h_bytes = 8.0 * (double) nout;
	}
	cudaFreeAsync(d_buf, s);
	return rc;
}

extern "C" int svtgpu_matmul_dev(svtgpu_matrix *m, const void *d_d_rowmajor,
				 int d_type, int64_t K, double *d_ans_rowmajor,
				 void *stream)
{
	SVT_CHECK(svtgpu_require_device());
	SVT_ARG(m != NULL && d_ans_rowmajor != NULL,
		"svtgpu_matmul_dev: NULL argument");
	SVT_ARG(d_type == SVTGPU_DOUBLE,
		"svtgpu_matmul_dev: the device operand must be double");
	cudaStream_t s = (cudaStream_t) stream;
	if (m->nrow == 0 || K == 0)
		return SVTGPU_OK;
	SVT_CUDA(cudaMemsetAsync(d_ans_rowmajor, 0,
				 8 * (size_t) (m->nrow * K), s));
	if (m->nnz == 0)
		return SVTGPU_OK;
	/* device form assumes clean operands: NA flags are not tracked */
	void *scratch = NULL;
	SVT_CHECK(svtgpu_scratch(m, 4 * (size_t) m->nrow + 64, &scratch));
	int32_t *row_na = (int32_t *) scratch;
	const double *D = (const double *) d_d_rowmajor;
	if (!(m->flags & SVTGPU_HAS_VALS))
		return launch_scatter<int32_t, true>(m, D, K, false,
				d_ans_rowmajor, row_na, NULL, s);
	if (svt_is_double(m->val_type))
		return launch_scatter<double, false>(m, D, K, false,
				d_ans_rowmajor, row_na, NULL, s);
	return launch_scatter<int32_t, false>(m, D, K, false, d_ans_rowmajor,
					      row_na, NULL, s);
}

extern "C" int svtgpu_matmul(svtgpu_matrix *m, const void *d, int d_type,
			     int64_t K, double *ans)
{
	SVT_CHECK(svtgpu_require_device());
	SVT_ARG(m != NULL && ans != NULL, "svtgpu_matmul: NULL argument");
	SVT_CHECK(check_product_types(m, d_type, "%*%"));
	SVT_CHECK(svtgpu_matrix_finish_upload(m));
	m->tm.kernel_ms = m->tm.d2h_ms = 0.0;
	m->tm.d2h_bytes = 0.0;
	m->tm.launches = 0;
	const int64_t nrow = m->nrow, n = m->nleaf;
	const size_t nout = (size_t) (nrow * K);
	if (nout == 0)
		return SVTGPU_OK;
	if (m->nnz == 0) {   /* t(x)@SVT == NULL: src/SparseMatrix_mult.c:389-390 */
		memset(ans, 0, sizeof(double) * nout);
		return SVTGPU_OK;
	}
	cudaStream_t s = 0;
	const size_t esz = d_type == SVTGPU_DOUBLE ? 8 : 4;
	const size_t raw_bytes = (esz * (size_t) (n * K) + 255) & ~(size_t) 255;
	const size_t rm_bytes = (8 * (size_t) (n * K) + 255) & ~(size_t) 255;
	const size_t info_bytes = (sizeof(SvtDenseColInfo) * (size_t) K + 255) &
				  ~(size_t) 255;
	const size_t prod_bytes = (8 * nout + 255) & ~(size_t) 255;
	const size_t na_bytes = (4 * (size_t) nrow + 255) & ~(size_t) 255;
	char *d_buf = NULL;
	SVT_CUDA(cudaMallocAsync((void **) &d_buf, raw_bytes + rm_bytes +
				 info_bytes + 2 * prod_bytes + na_bytes +
				 4 * nout + 256, s));
	char *p = d_buf;
	void *d_raw = p; p += raw_bytes;
	double *d_rm = (double *) p; p += rm_bytes;
	SvtDenseColInfo *d_info = (SvtDenseColInfo *) p; p += info_bytes;
	double *d_prod = (double *) p; p += prod_bytes;
	double *d_ans = (double *) p; p += prod_bytes;
	int32_t *d_row_na = (int32_t *) p; p += na_bytes;
	int32_t *d_hits = (int32_t *) p;
	int any_bad = 0;
	int64_t l0 = svtgpu_launch_count();
	int rc = SVTGPU_OK;
	if (n > 0)
		rc = prepare_dense(d, d_type, n, K, 0, n, K, d_rm, d_raw,
				   d_info, &any_bad, s);
	else
		SVT_CUDA(cudaMemsetAsync(d_info, 0, info_bytes, s));
	SvtTimer t;
	if (rc == SVTGPU_OK)
		rc = svt_timer_begin(&t, s);
	if (rc == SVTGPU_OK) {
		cudaError_t e = cudaMemsetAsync(d_prod, 0, prod_bytes, s);
		if (e == cudaSuccess)
			e = cudaMemsetAsync(d_row_na, 0, na_bytes, s);
		if (e == cudaSuccess && any_bad)
			e = cudaMemsetAsync(d_hits, 0, 4 * nout, s);
		if (e != cudaSuccess)
			rc = svtgpu_cuda_fail(e, "matmul memset", __FILE__,
					      __LINE__);
	}
	if (rc == SVTGPU_OK && m->nnz > 0) {
		int32_t *hits = any_bad ? d_hits : NULL;
		if (!(m->flags & SVTGPU_HAS_VALS))
			rc = launch_scatter<int32_t, true>(m, d_rm, K, any_bad,
					d_prod, d_row_na, hits, s);
		else if (svt_is_double(m->val_type))
			rc = launch_scatter<double, false>(m, d_rm, K, any_bad,
					d_prod, d_row_na, hits, s);
		else
			rc = launch_scatter<int32_t, false>(m, d_rm, K,
					any_bad, d_prod, d_row_na, hits, s);
	}
	if (rc == SVTGPU_OK) {
		matmul_finalize<<<grid_for((int64_t) nout, 256), 256, 0, s>>>(
			d_prod, d_row_na, any_bad ? d_hits : NULL, nrow, K,
			svt_is_double(m->val_type), d_info, d_ans, 0);
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess)
			rc = svtgpu_cuda_fail(e, "matmul_finalize", __FILE__,
					      __LINE__);
		svtgpu_count_launch(1);
	}
	int rc2 = svt_timer_end(&t, &m->tm.kernel_ms);
	if (rc == SVTGPU_OK)
		rc = rc2;
	m->tm.launches = (int) (svtgpu_launch_count() - l0);
	if (rc == SVTGPU_OK) {
		rc = svt_timer_begin(&t, s);
		cudaError_t e = cudaMemcpyAsync(ans, d_ans, 8 * nout,
						cudaMemcpyDeviceToHost, s);
		rc2 = svt_timer_end(&t, &m->tm.d2h_ms);
		if (e != cudaSuccess)
			rc = svtgpu_cuda_fail(e, "matmul D2H", __FILE__,
					      __LINE__);
		else if (rc == SVTGPU_OK)
			rc = rc2;
		m->tm.d2h_bytes = 8.0 * (double) nout;
	}
	cudaFreeAsync(d_buf, s);
	return rc;
}
