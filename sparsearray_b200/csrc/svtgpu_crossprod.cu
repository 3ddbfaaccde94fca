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
/* NA or NaN */
__device__ __forceinline__ bool val_is_special(int32_t x) { return x == SVT_NA_INT; }
__device__ __forceinline__ bool val_is_special(double x) { return svt_isnan(x); }

/* Leaf flags (SVT_LEAF_*) from the entries seen so far, in leaf order: `sp`
 * = this lane holds an NA/NaN entry, `na` = it is an NA; lanes are in entry
 * order.  *seen: an NA/NaN entry was met before this call. */
__device__ __forceinline__ int leaf_flag_update(int flag, bool *seen, bool sp,
						bool na)
{
	const unsigned msk = __ballot_sync(SVT_FULL_MASK, sp);
	if (msk == 0)
		return flag;
	const unsigned na_msk = __ballot_sync(SVT_FULL_MASK, sp && na);
	if (na_msk != 0)
		flag |= SVT_LEAF_HAS_NA;
	if (!*seen) {
		const int f = __ffs(msk) - 1;
		if (!((na_msk >> f) & 1u))
			flag |= SVT_LEAF_NAN_FIRST;
		*seen = true;
	}
	return flag;
}

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
			bool seen = false;
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
						my_v = (double) x;
					}
				}
				if (!LACUNAR) {
					const T x = e < end ? vals[e] : (T) 0;
					leaf_na = leaf_flag_update(leaf_na, &seen,
						e < end && val_is_special(x),
						val_is_na(x));
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
	       unsigned long long *__restrict__ row_first,
	       /* nrow or NULL: min over the row's NA/NaN entries of
		  (leaf << 1 | is-NaN), initialised to all ones */
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
						if (k0 == 0 && val_is_special(x)) {
							if (val_is_na(x))
								row_na[my_off] = 1;
							if (row_first != NULL)
								atomicMin(row_first +
								    my_off,
								    ((unsigned long long)
								     leaf << 1) |
								    (val_is_na(x)
								     ? 0ull : 1ull));
						}
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
		const unsigned long long *__restrict__ row_first,
		const int32_t *__restrict__ hits, int64_t nrow, int64_t K,
		int is_double, const SvtDenseColInfo *__restrict__ info,
		double *__restrict__ ans, int ans_rowmajor)
{
	const int64_t total = nrow * K;
	const int64_t stride = (int64_t) gridDim.x * blockDim.x;
	for (int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	     t < total; t += stride) {
		const int64_t i = t / K, k = t - i * K;
		int flag = row_na[i] ? SVT_LEAF_HAS_NA : 0;
		if (row_first != NULL && row_first[i] != ~0ull &&
		    (row_first[i] & 1ull))
			flag |= SVT_LEAF_NAN_FIRST;
		const double v = svt_dot_finalize(is_double, prod[t], flag,
				hits != NULL ? hits[t] : 0, info[k]);
		if (ans_rowmajor) ans[t] = v;
		else              ans[i + k * nrow] = v;
	}
}

/* ------------------------------------------------------------------------
 * crossprod_strips: the gather at shared-memory speed.
 *
 * The dense operand does not fit shared memory, but a slab of `strip_rows`
 * of its rows does.  A CTA owns a chunk of leaves (balanced by nonzeros) and
 * walks the row strips one after the other: it loads the slab of strip s,
 * then its warps take the chunk's leaves in turn and, for each, gather-
 * multiply the part of the leaf that falls into the strip (one contiguous
 * sub-run: offsets ascend; split points from the cached row_split table) and
 * add the K partial sums into the leaf's row of a row-major [nleaf][K]
 * result.  A leaf is always handled by the same warp, so these updates are
 * plain loads and stores that stay in L2.  Per nonzero the K dense values now
 * come out of shared memory instead of L2.
 *
 * Inside a warp the two half-warps work on two different nonzeros; lane c of
 * a half owns dense columns 4c .. 4c+3 (two 16-byte shared loads, four FMAs),
 * so K <= 64.  (offset, value) pairs are streamed into a register ring CP_D
 * leaves ahead, as in row_strips.
 */
#define CP_D 4   /* leaves prefetched ahead per warp */
#define CP_U 2   /* 32-wide slots per sub-run held in the ring */

struct CpStripParams {
	const int32_t *offs;
	const void *vals;
	const int64_t *leaf_ptr;
	const int32_t *split;
	int64_t nleaf, nnz, nrow;
	int nchunks, nstrips, strip_rows;
	int K, KP;                 /* columns, padded pitch (multiple of 4) */
	const double *Y;           /* [nrow][K] row-major */
	double *out;               /* [nparts][nleaf][K] row-major */
	int32_t *leaf_na;          /* [nparts][nleaf] */
	int slab_mode;             /* panels: one range of slabs per CTA */
};

template <typename T>
__device__ __forceinline__ T cp_ldg(const T *p) { return __ldg(p); }

template <typename T, bool LACUNAR, bool WIDE>   /* WIDE: more than 32 columns */
__global__ void __launch_bounds__(512, 1)
crossprod_strips(CpStripParams P)
{
	extern __shared__ __align__(16) double Ys[];   /* strip_rows x KP */
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int W = blockDim.x >> 5;
	const int half = lane >> 4;           /* which nonzero of a pair */
	/* lane c of a half owns columns 2c, 2c+1 and 32+2c, 32+2c+1: each
	   16-byte load of a half-warp covers one contiguous 256-byte piece of
	   a dense row (no bank conflicts inside a half) */
	const int c2 = (lane & 15) * 2;
	const bool on_a = c2 < P.KP, on_b = 32 + c2 < P.KP;
	const T *vals = (const T *) P.vals;
	const int K = P.K, KP = P.KP;

	int64_t l0, l1;
	{
		int64_t bounds[2];
		for (int k = 0; k < 2; k++) {
			const int c = (int) blockIdx.x + k;
			if (c >= P.nchunks) { bounds[k] = P.nleaf; continue; }
			const int64_t target = (int64_t) ((double) P.nnz *
					((double) c / (double) P.nchunks));
			int64_t lo = 0, hi = P.nleaf;
			while (lo < hi) {
				int64_t mid = lo + ((hi - lo) >> 1);
				if (P.leaf_ptr[mid] < target) lo = mid + 1;
				else                          hi = mid;
			}
			bounds[k] = c == 0 ? 0 : lo;
		}
		l0 = bounds[0];
		l1 = bounds[1];
	}
	/* this warp's leaves: l0 + warp + j * W, j = 0 .. nj - 1 */
	const int64_t nj = l1 - l0 > warp ? (l1 - l0 - warp + W - 1) / W : 0;

	for (int s = 0; s < P.nstrips; s++) {
		const int row0 = s * P.strip_rows;
		int rows = (int) (P.nrow - row0 < P.strip_rows ? P.nrow - row0
								: P.strip_rows);
		if (rows < 0) rows = 0;
		__syncthreads();
		/* slab -> shared memory: a warp per row, lanes over columns,
		   four rows in flight per warp */
		for (int r = warp; r < rows; r += 4 * W) {
			double t[4][2];
#pragma unroll
			for (int q = 0; q < 4; q++) {
				const int rr = r + q * W;
				const double *src = P.Y + (size_t) (row0 + rr) * K;
				t[q][0] = rr < rows && lane < K ? src[lane] : 0.0;
				t[q][1] = rr < rows && lane + 32 < K
					? src[lane + 32] : 0.0;
			}
#pragma unroll
			for (int q = 0; q < 4; q++) {
				const int rr = r + q * W;
				if (rr < rows) {
					if (lane < KP)
						Ys[rr * KP + lane] = t[q][0];
					if (lane + 32 < KP)
						Ys[rr * KP + lane + 32] = t[q][1];
				}
			}
		}
		__syncthreads();
		/* 32-bit shared addresses of this lane's two column pairs in
		   the (virtual) row 0 of the slab.  Lanes whose second pair
		   lies beyond the padded width read lane 0's pair instead (a
		   broadcast: no extra wavefront) and their sums are never
		   stored -- so loads and FMAs carry no predicates. */
		const uint32_t kp_bytes = (uint32_t) KP * 8u;
		const uint32_t ys0 = (uint32_t) __cvta_generic_to_shared(Ys) -
				     (uint32_t) row0 * kp_bytes;
		/* (a 16-byte shared load is served 8 lanes at a time: an idle
		   lane copies the first lane of its own group of 8 when that
		   one is active, so it adds no wavefront and no conflict) */
		const int q2 = (lane & 8) * 2;     /* first column pair of the group */
		const uint32_t ya_s = ys0 + (uint32_t)
			(on_a ? c2 : (q2 < KP ? q2 : 0)) * 8u;
		const uint32_t yb_s = ys0 + (uint32_t)
			(on_b ? 32 + c2 : (32 + q2 < KP ? 32 + q2 : 32)) * 8u;

		auto subrun = [&](int64_t j, int64_t &lo, int &n) {
			lo = 0; n = 0;
			if (j < nj) {
				const int64_t leaf = l0 + warp + j * W;
				const int64_t start = P.leaf_ptr[leaf];
				const int nz = (int) (P.leaf_ptr[leaf + 1] - start);
				const int a = s == 0 ? 0
					: P.split[(int64_t) (s - 1) * P.nleaf + leaf];
				const int b = s == P.nstrips - 1 ? nz
					: P.split[(int64_t) s * P.nleaf + leaf];
				lo = start + a;
				n = b - a;
			}
		};
		int32_t boff[CP_D][CP_U];
		T bval[CP_D][CP_U];
		int64_t blo[CP_D];
		int bn[CP_D];
		auto fetch = [&](int d, int64_t lo, int n, int64_t j) {
			blo[d] = lo;
			bn[d] = n;
			/* the leaf's result row is read-modify-written when its
			   sub-run has been applied, CP_D leaves from now: start
			   bringing it in (the exposed miss cost 9 % of the
			   kernel; keeping a group of leaves' rows hot in L2 by
			   looping slabs inside leaf groups was tried and lost:
			   74.9 vs 68.9 ms) */
			if (n > 0 && half == 0) {
				const double *o = P.out +
					(size_t) (l0 + warp + j * W) * K + c2;
				if (c2 < K)
					asm volatile("prefetch.global.L1 [%0];"
						     :: "l"(o));
				if (c2 + 32 < K)
					asm volatile("prefetch.global.L1 [%0];"
						     :: "l"(o + 32));
			}
#pragma unroll
			for (int k = 0; k < CP_U; k++) {
				const int e = k * 32 + lane;
				boff[d][k] = row0;
				bval[d][k] = (T) 0;
				if (e < n) {
					boff[d][k] = cp_ldg(P.offs + lo + e);
					if (!LACUNAR)
						bval[d][k] = cp_ldg(vals + lo + e);
				}
			}
		};
		/* one pair of nonzeros (one per half-warp) */
		auto fma4 = [&](int off, double v, double &s0, double &s1,
				double &s2, double &s3) {
			const uint32_t r = (uint32_t) off * kp_bytes;
			double ax, ay;
			asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
				     : "=d"(ax), "=d"(ay) : "r"(ya_s + r));
			s0 += v * ax; s1 += v * ay;
			if (WIDE) {
				double bx, by;
				asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
					     : "=d"(bx), "=d"(by) : "r"(yb_s + r));
				s2 += v * bx; s3 += v * by;
			}
		};
		auto apply = [&](int d, int64_t j) {
			const int n = bn[d];
			if (n == 0)
				return;
			const int64_t leaf = l0 + warp + j * W;
			/* the leaf's result row so far: loaded now, needed only
			   after the gather loop (the read-modify-write at the
			   end of a sub-run was the kernel's largest stall) */
			double *const orow = P.out + (size_t) leaf * K + c2;
			double o0 = 0.0, o1 = 0.0, o2 = 0.0, o3 = 0.0;
			if (half == 0) {
				/* (a row of an odd K is only 8-byte aligned) */
				if (c2 + 1 < K && (K & 1) == 0) {
					asm volatile("ld.global.v2.f64 {%0, %1}, [%2];"
						     : "=d"(o0), "=d"(o1) : "l"(orow));
				} else if (c2 < K) {
					asm volatile("ld.global.f64 %0, [%1];"
						     : "=d"(o0) : "l"(orow));
					if (c2 + 1 < K)
						asm volatile("ld.global.f64 %0, [%1];"
							     : "=d"(o1) : "l"(orow + 1));
				}
				if (WIDE && c2 + 33 < K && (K & 1) == 0) {
					asm volatile("ld.global.v2.f64 {%0, %1}, [%2];"
						     : "=d"(o2), "=d"(o3)
						     : "l"(orow + 32));
				} else if (WIDE && c2 + 32 < K) {
					asm volatile("ld.global.f64 %0, [%1];"
						     : "=d"(o2) : "l"(orow + 32));
					if (c2 + 33 < K)
						asm volatile("ld.global.f64 %0, [%1];"
							     : "=d"(o3) : "l"(orow + 33));
				}
			}
			double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
			/* SVT_LEAF_* flags of this sub-run, entries in order */
			int flag = 0;
			bool seen = false;
#pragma unroll
			for (int k = 0; k < CP_U; k++) {
				int cnt = n - k * 32;
				if (cnt <= 0)
					break;
				if (cnt > 32) cnt = 32;
				if (!LACUNAR)
					flag = leaf_flag_update(flag, &seen,
						lane < cnt &&
						val_is_special(bval[d][k]),
						val_is_na(bval[d][k]));
				const double myv = LACUNAR
					? (lane < cnt ? 1.0 : 0.0)
					: (double) bval[d][k];
				/* padding lanes carry (row0, 0): harmless */
#pragma unroll 4
				for (int i = 0; i < cnt; i += 2) {
					const int off = __shfl_sync(SVT_FULL_MASK,
						boff[d][k], i + half);
					const double v = __shfl_sync(SVT_FULL_MASK,
						myv, i + half);
					fma4(off, v, s0, s1, s2, s3);
				}
			}
			/* sub-runs longer than the ring holds */
			for (int e0 = CP_U * 32; e0 < n; e0 += 2) {
				const int e = e0 + half;
				int off = row0;
				double v = 0.0;
				if (e < n) {
					off = P.offs[blo[d] + e];
					if (LACUNAR) {
						v = 1.0;
					} else {
						const T x = vals[blo[d] + e];
						v = (double) x;
					}
				}
				if (!LACUNAR) {
					/* lanes 0 and 16 speak for the two
					   entries of the pair, in order */
					const T x = e < n ? vals[blo[d] + e] : (T) 0;
					const bool lead = (lane & 15) == 0;
					flag = leaf_flag_update(flag, &seen,
						lead && e < n && val_is_special(x),
						val_is_na(x));
				}
				fma4(off, v, s0, s1, s2, s3);
			}
			/* the two halves hold the even / odd nonzeros */
			s0 += __shfl_xor_sync(SVT_FULL_MASK, s0, 16);
			s1 += __shfl_xor_sync(SVT_FULL_MASK, s1, 16);
			s2 += __shfl_xor_sync(SVT_FULL_MASK, s2, 16);
			s3 += __shfl_xor_sync(SVT_FULL_MASK, s3, 16);
			if (half == 0) {
				if (c2 + 0 < K) orow[0] = o0 + s0;
				if (c2 + 1 < K) orow[1] = o1 + s1;
				if (c2 + 32 < K) orow[32] = o2 + s2;
				if (c2 + 33 < K) orow[33] = o3 + s3;
			}
			/* slabs are visited in ascending row order by the same
			   warp: bit 2 remembers that an earlier slab already
			   decided which NA/NaN entry of the leaf comes first */
			if (seen && lane == 0) {
				const int old = P.leaf_na[leaf];
				int nw = old | (flag & SVT_LEAF_HAS_NA) | 4;
				if (!(old & 4))
					nw |= flag & SVT_LEAF_NAN_FIRST;
				P.leaf_na[leaf] = nw;
			}
		};

		/* bounds of 32 of this warp's leaves per batch (lane i: j =
		   batch + i), fetched one batch ahead */
		int64_t cur_lo, nxt_lo;
		int cur_n, nxt_n;
		subrun(lane, cur_lo, cur_n);
		subrun(32 + lane, nxt_lo, nxt_n);
#pragma unroll
		for (int d = 0; d < CP_D; d++) {
			const int64_t lo = __shfl_sync(SVT_FULL_MASK, cur_lo, d);
			const int n = __shfl_sync(SVT_FULL_MASK, cur_n, d);
			fetch(d, lo, n, d);
		}
		for (int64_t jb = 0; jb < nj; jb += 32) {
			for (int i0 = 0; i0 < 32; i0 += CP_D) {
				if (jb + i0 >= nj)
					break;
#pragma unroll
				for (int d = 0; d < CP_D; d++) {
					const int i = i0 + d;
					apply(d, jb + i);
					const int jn = i + CP_D;
					int64_t lo;
					int n;
					if (jn < 32) {
						lo = __shfl_sync(SVT_FULL_MASK,
								 cur_lo, jn);
						n = __shfl_sync(SVT_FULL_MASK,
								cur_n, jn);
					} else {
						lo = __shfl_sync(SVT_FULL_MASK,
								 nxt_lo, jn - 32);
						n = __shfl_sync(SVT_FULL_MASK,
								nxt_n, jn - 32);
					}
					fetch(d, lo, n, jb + jn);
				}
			}
			cur_lo = nxt_lo;
			cur_n = nxt_n;
			subrun(jb + 64 + lane, nxt_lo, nxt_n);
		}
	}
}

/* ------------------------------------------------------------------------
 * crossprod_panels: the slab gather with the result kept ON CHIP.
 *
 * crossprod_strips walks the slabs once per CTA and read-modify-writes every
 * leaf's K partial sums through L2 at every slab (63 x 400 MB at the headline
 * size: 2.5x the algorithmic DRAM traffic), and broadcasts every nonzero to
 * the lanes with three shuffles -- which travel through the same
 * shared-memory pipe as the gather itself (ncu: that pipe 90 % busy).  Here:
 *
 *  - A CTA takes its leaves in PANELS of 512 (32 per warp) and walks all the
 *    slabs once per panel.  A panel's K x 512 partial sums live in tensor
 *    memory (TMEM: 256 KB per SM that this kernel has no other use for; a
 *    warp owns 32 leaves x 4 columns of its lane quarter), so the result is
 *    written to HBM exactly once and nothing is re-read.  The price is
 *    reloading the slabs once per panel out of L2 (nleaf / 512 x the dense
 *    operand: 26 GB at the headline size, one bulk async copy per slab).
 *  - A sub-run's nonzeros are staged as 16-byte records {shared address of
 *    the dense row, value} in a warp-private strip of shared memory; each
 *    half-warp then reads the record of its nonzero with ONE broadcast
 *    16-byte load (1 wavefront per pair of nonzeros instead of 3 shuffles).
 *  - The dense row is read as 16 + 8 + 8 bytes per lane (columns 2c, 2c+1 |
 *    32+c | 48+c for lane c of a half): K = 50 costs 3.5 wavefronts per
 *    nonzero instead of 4.
 *  - No per-leaf state in memory: positions, NA flags and the bounds of the
 *    sub-runs of a warp's 32 leaves sit in the registers of lane i.
 */
#define CPN_D 4
#define CPN_U 2
#define CPN_REC_BYTES (CPN_U * 32 * 16)   /* records of one warp */
#define CPN_HDR 64                         /* TMEM slot, mbarrier */

__device__ __forceinline__ void tmem_alloc512(uint32_t smem_dst)
{
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 "
		     "[%0], 512;" :: "r"(smem_dst) : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"
		     ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc512(uint32_t taddr)
{
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;"
		     :: "r"(taddr) : "memory");
}

__device__ __forceinline__ void tmem_fence_before(void)
{
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void tmem_fence_after(void)
{
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

/* 4 consecutive 32-bit columns of this thread's TMEM lane */
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t &a,
					 uint32_t &b, uint32_t &c, uint32_t &d)
{
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
		     : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr) : "memory");
}

/* the loaded registers are valid after this (they pass through the
   statement so that nothing that reads them is scheduled above it) */
__device__ __forceinline__ void tmem_wait_ld4(uint32_t &a, uint32_t &b,
					      uint32_t &c, uint32_t &d)
{
	asm volatile("tcgen05.wait::ld.sync.aligned;"
		     : "+r"(a), "+r"(b), "+r"(c), "+r"(d) :: "memory");
}

__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b,
					 uint32_t c, uint32_t d)
{
	asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
		     :: "r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void tmem_wait_st(void)
{
	asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

/* NL: 16-byte + 8-byte + 8-byte loads per lane and nonzero that K needs
   (1: K <= 32, 2: K <= 48, 3: K <= 64).  ACC_TMEM: partial sums in tensor
   memory, else read-modify-written in the (zero-initialised) result rows,
   which a panel keeps hot in L2.  BULK: the slab is one contiguous piece of
   the dense operand (even K): bulk async copy.

   Two ways of cutting the work into one piece per CTA (P.slab_mode):
     leaves  a chunk of leaves (balanced by nonzeros) x all slabs -- many
             leaves, few slabs: crossprod(svt, Y);
     slabs   all leaves x a range of slabs, the piece's sums going to its own
             partial result (summed in a fixed order by crossprod_finish) --
             few long leaves, many slabs: `svt %*% D` on the transpose. */
template <typename T, bool LACUNAR, int NL, bool ACC_TMEM, bool BULK>
__global__ void __launch_bounds__(512, 1)
crossprod_panels(CpStripParams P)
{
	/* NL = 16-column chunks of a dense row read with 16-byte loads (1..4);
	   NL == 5 stands for 3 chunks + the two columns 48, 49 read with one
	   8-byte load (K = 49, 50).  A QUARTER-warp works on one nonzero: lane j
	   of a quarter reads columns 16m + 2j, 16m + 2j + 1 of chunk m, so every
	   16-byte load instruction of a warp fetches 4 x 128 contiguous bytes
	   (4 nonzeros, no bank conflicts, 4 cycles of the shared-memory pipe);
	   measured costs of the alternatives: tools/microbench/lds_patterns.cu */
	constexpr int NC = NL == 5 ? 3 : NL;
	constexpr bool TAIL = NL == 5;
	extern __shared__ __align__(128) unsigned char cpn_smem[];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int W = blockDim.x >> 5;
	const int q = lane >> 3;
	const int j = lane & 7;
	const T *vals = (const T *) P.vals;
	const int K = P.K, KP = P.KP;
	const uint32_t pitch = (uint32_t) KP * 8u;
	uint32_t *tm_slot = (uint32_t *) cpn_smem;
	const uint32_t bar = svt_smem_u32(cpn_smem + 16);
	const uint32_t rec_s = svt_smem_u32(cpn_smem + CPN_HDR) +
			       (uint32_t) warp * CPN_REC_BYTES;
	double *Ys = (double *) (cpn_smem + CPN_HDR + (size_t) W * CPN_REC_BYTES);
	const uint32_t ys_s = svt_smem_u32(Ys);
	/* byte offsets of this lane's columns inside a dense row; a lane whose
	   columns lie beyond the (even) padded width re-reads the first pair of
	   its chunk: a broadcast inside its own quarter (no conflict, no extra
	   cycle), and its sums are never stored */
	uint32_t offm[4];
#pragma unroll
	for (int m = 0; m < 4; m++)
		offm[m] = (uint32_t) (16 * m + 2 * j < KP ? 16 * m + 2 * j
							    : 16 * m) * 8u;
	const uint32_t offT = (uint32_t) (48 + (j & 1)) * 8u;
	/* after the sums of the four quarters have been combined, quarter m
	   holds chunk m: columns 16m + 2j, 16m + 2j + 1 (quarter 3 of the
	   K = 50 layout: column 48 + j, lanes j < 2) */
	const int colx = TAIL && q == 3 ? (j < 2 ? 48 + j : KP) : 16 * q + 2 * j;
	const int coly = TAIL && q == 3 ? KP : 16 * q + 2 * j + 1;

	int64_t l0 = 0, l1 = P.nleaf;
	int s0 = 0, s1 = P.nstrips;
	double *out = P.out;
	int32_t *leaf_na = P.leaf_na;
	if (P.slab_mode) {
		s0 = (int) (((int64_t) P.nstrips * blockIdx.x) / P.nchunks);
		s1 = (int) (((int64_t) P.nstrips * (blockIdx.x + 1)) / P.nchunks);
		out += (size_t) blockIdx.x * (size_t) P.nleaf * K;
		leaf_na += (size_t) blockIdx.x * (size_t) P.nleaf;
	} else {
		int64_t bounds[2];
		for (int k = 0; k < 2; k++) {
			const int ch = (int) blockIdx.x + k;
			if (ch >= P.nchunks) { bounds[k] = P.nleaf; continue; }
			const int64_t target = (int64_t) ((double) P.nnz *
					((double) ch / (double) P.nchunks));
			int64_t lo = 0, hi = P.nleaf;
			while (lo < hi) {
				int64_t mid = lo + ((hi - lo) >> 1);
				if (P.leaf_ptr[mid] < target) lo = mid + 1;
				else                          hi = mid;
			}
			bounds[k] = ch == 0 ? 0 : lo;
		}
		l0 = bounds[0];
		l1 = bounds[1];
	}
	/* panels of equal size: pw leaves per warp (<= 32) */
	int pw = 32;
	{
		const int64_t L = l1 - l0;
		const int64_t np = (L + 32 * W - 1) / (32 * W);
		if (np > 0) {
			const int64_t per = (L + np - 1) / np;
			pw = (int) ((per + W - 1) / W);
		}
	}

	uint32_t tcol0 = 0, tbase = 0;
	if (BULK && threadIdx.x == 0) {
		svt_mbar_init(bar, 1);
		svt_mbar_init_fence();
	}
	if (ACC_TMEM) {
		if (warp == 0)
			tmem_alloc512(svt_smem_u32(tm_slot));
		tmem_fence_before();
	}
	__syncthreads();
	if (ACC_TMEM) {
		tmem_fence_after();
		tbase = *tm_slot;
		/* a warp reaches the 32 TMEM lanes of its quarter (warp % 4);
		   the four warps of a quarter take 128 columns each */
		tcol0 = tbase + ((uint32_t) (warp & 3) << 21) +
			(uint32_t) (warp >> 2) * 128u;
	}
	uint32_t phase = 0;

	for (int64_t pb = l0; pb < l1; pb += (int64_t) W * pw) {
		const int64_t wleaf0 = pb + (int64_t) warp * pw;
		const int64_t myleaf = wleaf0 + lane;
		const bool have = lane < pw && myleaf < l1;
		int64_t start_l = 0;
		int nz_l = 0;
		if (have) {
			start_l = P.leaf_ptr[myleaf];
			nz_l = (int) (P.leaf_ptr[myleaf + 1] - start_l);
		}
		/* the warp's leaves are consecutive: their nonzeros are one
		   contiguous piece, addressed relative to its first entry */
		const int64_t wstart = __shfl_sync(SVT_FULL_MASK, start_l, 0);
		const uint32_t rel_l = have ? (uint32_t) (start_l - wstart) : 0u;
		const int32_t *woffs = P.offs + wstart;
		const T *wvals = LACUNAR ? NULL : vals + wstart;
		int lflag = 0;            /* my leaf's SVT_LEAF_* flags */
		bool lseen = false;
		if (ACC_TMEM) {
#pragma unroll 4
			for (int i = 0; i < 32; i++)
				tmem_st4(tcol0 + 4u * (uint32_t) i, 0u, 0u, 0u, 0u);
		}
		/* my leaf's entries: [a_l, b_l) fall into the current slab,
		   [b_l, b_nx) into the next one */
		auto split_at = [&](int s) -> int {   /* entries in slabs < s */
			if (!have || s <= 0) return 0;
			if (s >= P.nstrips) return nz_l;
			return cp_ldg(P.split + (int64_t) (s - 1) * P.nleaf + myleaf);
		};
		int a_l = split_at(s0);
		int b_l = split_at(s0 + 1);
		if (s0 >= s1) b_l = a_l;
		uint32_t lo_l = rel_l + (uint32_t) a_l;
		int n_l = b_l - a_l;
		uint32_t lo_nx = 0;
		int n_nx = 0;

		int32_t boff[CPN_D][CPN_U];
		T bval[CPN_D][CPN_U];
		uint32_t blo[CPN_D];
		int bn[CPN_D];
		/* sub-run i of the current slab (i < 32) or i - 32 of the next
		   one: the ring runs on across the slab boundary */
		auto fetch = [&](int d, int i) {
			const uint32_t lo = __shfl_sync(SVT_FULL_MASK,
					i < 32 ? lo_l : lo_nx, i & 31);
			const int n = __shfl_sync(SVT_FULL_MASK,
					i < 32 ? n_l : n_nx, i & 31);
			blo[d] = lo;
			bn[d] = n;
#pragma unroll
			for (int k = 0; k < CPN_U; k++) {
				const int e = k * 32 + lane;
				boff[d][k] = 0;
				bval[d][k] = (T) 0;
				if (e < n) {
					boff[d][k] = cp_ldg(woffs + lo + e);
					if (!LACUNAR)
						bval[d][k] = cp_ldg(wvals + lo + e);
				}
			}
		};
#pragma unroll
		for (int d = 0; d < CPN_D; d++)
			fetch(d, d);

		for (int s = s0; s < s1; s++) {
			/* the bounds of the next slab's sub-runs, needed when the
			   ring runs ahead into it */
			const int b_nx = s + 1 < s1 ? split_at(s + 2) : b_l;
			lo_nx = rel_l + (uint32_t) b_l;
			n_nx = b_nx - b_l;
			const int row0 = s * P.strip_rows;
			int rows = (int) (P.nrow - row0 < P.strip_rows
					  ? P.nrow - row0 : P.strip_rows);
			if (rows < 0) rows = 0;
			__syncthreads();          /* previous slab no longer read */
			if (BULK) {
				if (threadIdx.x == 0) {
					const uint32_t bytes = (uint32_t) rows * pitch;
					const char *src = (const char *) P.Y +
						(size_t) row0 * pitch;
					svt_mbar_arrive_expect_tx(bar, bytes);
					for (uint32_t o = 0; o < bytes; o += 32768u)
						svt_bulk_g2s(ys_s + o, src + o,
							bytes - o < 32768u
							? bytes - o : 32768u, bar);
				}
				svt_mbar_wait(bar, phase);
				phase ^= 1u;
			} else {
				for (int r = warp; r < rows; r += 4 * W) {
					double t[4][2];
#pragma unroll
					for (int q = 0; q < 4; q++) {
						const int rr = r + q * W;
						const double *src = P.Y +
							(size_t) (row0 + rr) * K;
						t[q][0] = rr < rows && lane < K
							? src[lane] : 0.0;
						t[q][1] = rr < rows && lane + 32 < K
							? src[lane + 32] : 0.0;
					}
#pragma unroll
					for (int q = 0; q < 4; q++) {
						const int rr = r + q * W;
						if (rr < rows) {
							if (lane < KP)
								Ys[rr * KP + lane] = t[q][0];
							if (lane + 32 < KP)
								Ys[rr * KP + lane + 32] = t[q][1];
						}
					}
				}
				__syncthreads();
			}
			/* shared address of the (virtual) row 0 of the operand */
			const uint32_t ys0 = ys_s - (uint32_t) row0 * pitch;

			auto apply = [&](int d, int i) {
				const int n = bn[d];
				if (n == 0)
					return;
				const uint32_t taddr = tcol0 + 4u * (uint32_t) i;
				uint32_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;
				double *const orow = out + (size_t) (wleaf0 + i) * K;
				double ox = 0.0, oy = 0.0;
				if (ACC_TMEM) {
					tmem_wait_st();
					tmem_ld4(taddr, t0, t1, t2, t3);
				} else {
					if (colx < K) ox = orow[colx];
					if (coly < K) oy = orow[coly];
				}
				double acc[4][2];
#pragma unroll
				for (int m = 0; m < 4; m++)
					acc[m][0] = acc[m][1] = 0.0;
				int flag = 0;
				bool seen = false;
				for (int done = 0; done < n; done += CPN_U * 32) {
					const int cnt = n - done < CPN_U * 32
						? n - done : CPN_U * 32;
					const int cnt2 = (cnt + 3) & ~3;
					if (done > 0)
						__syncwarp();
#pragma unroll
					for (int k = 0; k < CPN_U; k++) {
						const int e = k * 32 + lane;
						int32_t off = boff[d][k];
						T x = bval[d][k];
						if (done > 0) {   /* beyond the ring: rare */
							off = 0;
							x = (T) 0;
							if (e < cnt) {
								off = woffs[blo[d] + done + e];
								if (!LACUNAR)
									x = wvals[blo[d] + done + e];
							}
						}
						if (k * 32 < cnt2 && e < cnt2) {
							const double v = e < cnt
								? (LACUNAR ? 1.0 : (double) x)
								: 0.0;
							const uint32_t ra = e < cnt
								? ys0 + (uint32_t) off * pitch
								: ys_s;
							asm volatile(
							    "st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
							    :: "r"(rec_s + (uint32_t) e * 16u),
							       "r"(ra), "r"(0),
							       "r"(__double2loint(v)),
							       "r"(__double2hiint(v))
							    : "memory");
						}
						if (!LACUNAR && k * 32 < cnt)
							flag = leaf_flag_update(flag, &seen,
								e < cnt && val_is_special(x),
								val_is_na(x));
					}
					__syncwarp();
					/* four nonzeros per trip, one per quarter-warp */
#pragma unroll 2
					for (int i2 = 0; i2 < cnt; i2 += 4) {
						uint32_t ra, pad, vlo, vhi;
						asm volatile(
						    "ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
						    : "=r"(ra), "=r"(pad), "=r"(vlo), "=r"(vhi)
						    : "r"(rec_s + (uint32_t) (i2 + q) * 16u));
						const double v = __hiloint2double((int) vhi,
										  (int) vlo);
#pragma unroll
						for (int m = 0; m < NC; m++) {
							double ax, ay;
							asm volatile(
							    "ld.shared.v2.f64 {%0, %1}, [%2];"
							    : "=d"(ax), "=d"(ay)
							    : "r"(ra + offm[m]));
							acc[m][0] += v * ax;
							acc[m][1] += v * ay;
						}
						if (TAIL) {
							double tx;
							asm volatile("ld.shared.f64 %0, [%1];"
								     : "=d"(tx) : "r"(ra + offT));
							acc[3][0] += v * tx;
						}
					}
				}
				/* combine the four quarters so that quarter m ends up
				   with chunk m: swap halves of the chunk list with the
				   quarter 2 away, then with the neighbour */
				double mx, my;
				{
					const bool hi = q >= 2;
					const double k0x = (hi ? acc[2][0] : acc[0][0]) +
						__shfl_xor_sync(SVT_FULL_MASK,
							hi ? acc[0][0] : acc[2][0], 16);
					const double k0y = (hi ? acc[2][1] : acc[0][1]) +
						__shfl_xor_sync(SVT_FULL_MASK,
							hi ? acc[0][1] : acc[2][1], 16);
					const double k1x = (hi ? acc[3][0] : acc[1][0]) +
						__shfl_xor_sync(SVT_FULL_MASK,
							hi ? acc[1][0] : acc[3][0], 16);
					const double k1y = (hi ? acc[3][1] : acc[1][1]) +
						__shfl_xor_sync(SVT_FULL_MASK,
							hi ? acc[1][1] : acc[3][1], 16);
					const bool odd = q & 1;
					mx = (odd ? k1x : k0x) + __shfl_xor_sync(
						SVT_FULL_MASK, odd ? k0x : k1x, 8);
					my = (odd ? k1y : k0y) + __shfl_xor_sync(
						SVT_FULL_MASK, odd ? k0y : k1y, 8);
				}
				if (ACC_TMEM) {
					tmem_wait_ld4(t0, t1, t2, t3);
					const double ax = __hiloint2double((int) t1, (int) t0) + mx;
					const double ay = __hiloint2double((int) t3, (int) t2) + my;
					tmem_st4(taddr, (uint32_t) __double2loint(ax),
						 (uint32_t) __double2hiint(ax),
						 (uint32_t) __double2loint(ay),
						 (uint32_t) __double2hiint(ay));
				} else {
					if (colx < K) orow[colx] = ox + mx;
					if (coly < K) orow[coly] = oy + my;
				}
				/* slabs come in ascending row order: the first NA/NaN
				   entry of the leaf is the first one ever seen */
				if (seen && lane == i) {
					lflag |= flag & SVT_LEAF_HAS_NA;
					if (!lseen)
						lflag |= flag & SVT_LEAF_NAN_FIRST;
					lseen = true;
				}
				__syncwarp();     /* the records may be overwritten */
			};

			for (int i0 = 0; i0 < 32; i0 += CPN_D) {
#pragma unroll
				for (int d = 0; d < CPN_D; d++) {
					apply(d, i0 + d);
					fetch(d, i0 + d + CPN_D);
				}
			}
			a_l = b_l;
			b_l = b_nx;
			lo_l = lo_nx;
			n_l = n_nx;
		}

		/* the panel is complete: its rows go to HBM once */
		if (ACC_TMEM) {
			tmem_wait_st();
			for (int i = 0; i < pw; i++) {
				if (wleaf0 + i >= l1)
					break;
				uint32_t t0, t1, t2, t3;
				tmem_ld4(tcol0 + 4u * (uint32_t) i, t0, t1, t2, t3);
				tmem_wait_ld4(t0, t1, t2, t3);
				double *const orow = out + (size_t) (wleaf0 + i) * K;
				if (colx < K)
					orow[colx] = __hiloint2double((int) t1, (int) t0);
				if (coly < K)
					orow[coly] = __hiloint2double((int) t3, (int) t2);
			}
		}
		/* bit 2: some NA / NaN entry was met (for merging slab pieces) */
		if (have)
			leaf_na[myleaf] = lflag | (lseen ? 4 : 0);
	}

	if (ACC_TMEM) {
		tmem_fence_before();
		__syncthreads();
		tmem_fence_after();
		if (warp == 0)
			tmem_dealloc512(tbase);
	}
}

/* row-major sums + leaf NA flags -> the answer in its final orientation.
   nparts > 1: the partial results of the slab pieces (ascending rows) are
   summed in that order, and the first piece that met an NA / NaN entry says
   which kind came first. */
__global__ void __launch_bounds__(256)
crossprod_finish(const double *__restrict__ rm,
		 const int32_t *__restrict__ leaf_na, int64_t nleaf, int64_t K,
		 int nparts, int is_double,
		 const SvtDenseColInfo *__restrict__ info,
		 double *__restrict__ ans, int svt_left)
{
	const int64_t total = nleaf * K;
	const int64_t stride = (int64_t) gridDim.x * blockDim.x;
	for (int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	     t < total; t += stride) {
		/* walk the OUTPUT linearly so the stores coalesce */
		int64_t l, k;
		if (svt_left) { k = t / nleaf; l = t - k * nleaf; }
		else          { l = t / K;     k = t - l * K; }
		double sum = rm[l * K + k];
		int flag = leaf_na[l];
		for (int p = 1; p < nparts; p++) {
			sum += rm[(size_t) p * (size_t) total + l * K + k];
			const int f = leaf_na[(size_t) p * (size_t) nleaf + l];
			flag |= f & SVT_LEAF_HAS_NA;
			if (!(flag & 4))
				flag |= f & SVT_LEAF_NAN_FIRST;
			flag |= f & 4;
		}
		ans[t] = svt_dot_finalize(is_double, sum, flag & 3, 0, info[k]);
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

struct CpPlan {
	int ok, nstrips, strip_rows, KP, nchunks, warps;
	int panels;          /* crossprod_panels (else crossprod_strips) */
	int acc_tmem, bulk;  /* panels: accumulators in TMEM; bulk slab copies */
	int slab_mode;       /* panels: a CTA owns a range of slabs, not leaves */
	int nparts;          /* partial results to sum (slab mode: nchunks) */
	size_t smem;
};

/* strips pay off when the slab holds enough rows for sub-runs of a few dozen
   nonzeros; otherwise (large K, tiny matrices) the L2 gather kernel is used */
CpPlan plan_crossprod_strips(const svtgpu_matrix *m, int64_t K)
{
	CpPlan p;
	memset(&p, 0, sizeof(p));
	const char *impl = svtgpu_env("SVTGPU_CP_IMPL", "auto");
	if (strcmp(impl, "gather") == 0 || K < 1 || K > 64 ||
	    m->nleaf < 1 || !(m->flags & SVTGPU_HAS_OFFS))
		return p;
	if ((((uintptr_t) m->d_offs) & 3) != 0)
		return p;
	/* the panel kernel (result rows on chip, record broadcast) unless asked
	   for the older one; it addresses a warp's 32 leaves with 32-bit
	   positions */
	p.panels = strcmp(impl, "strips") != 0 &&
		   m->nrow < ((int64_t) 1 << 26);
	p.acc_tmem = strcmp(svtgpu_env("SVTGPU_CP_ACC", "tmem"), "tmem") == 0;
	p.KP = p.panels ? (int) ((K + 1) & ~(int64_t) 1)
			: (int) ((K + 3) & ~(int64_t) 3);
	p.bulk = p.panels && p.KP == K &&
		 strcmp(svtgpu_env("SVTGPU_CP_BULK", "on"), "on") == 0;
	const size_t budget = (size_t) 227 * 1024 - 1024 - 256 -
		(p.panels ? CPN_HDR + 16 * CPN_REC_BYTES : 0);
	int64_t rows = (int64_t) (budget / ((size_t) p.KP * 8));
	rows = rows / 8 * 8;
	if (rows > m->nrow) rows = (m->nrow + 7) / 8 * 8;
	const double density = m->nrow > 0 && m->nleaf > 0
		? (double) m->nnz / ((double) m->nrow * (double) m->nleaf) : 0.0;
	/* the register ring holds sub-runs of up to CP_U x 32 nonzeros; longer
	   ones finish on a slow scalar path: keep the expected sub-run (mean +
	   3 sigma) inside the ring by using shorter slabs than would fit
	   (small K) */
	if (density * (double) rows > 44.0) {
		int64_t r = (int64_t) (44.0 / density) / 8 * 8;
		if (r < 64) r = 64;
		if (r < rows) rows = r;
	}
	/* expected nonzeros of a leaf inside one slab */
	if (rows < 64 || density * (double) rows < 8.0 ||
	    m->nnz < 4 * 1024 * 1024)
		if (strcmp(impl, "force") != 0 && strcmp(impl, "strips") != 0)
			return p;
	if (rows < 8)
		return p;
	p.strip_rows = (int) rows;
	p.nstrips = (int) ((m->nrow + rows - 1) / rows);
	if (p.nstrips < 1) p.nstrips = 1;
	p.smem = (size_t) rows * p.KP * 8;
	if (p.panels)   /* header + records, and slack behind the last row for
			   the lanes that read past the padded width */
		p.smem += CPN_HDR + 16 * CPN_REC_BYTES + 512;
	p.warps = 16;
	p.nchunks = svtgpu_sm_count();
	if ((int64_t) p.nchunks > m->nleaf)
		p.nchunks = (int) m->nleaf;
	p.nparts = 1;
	if (p.panels) {
		/* slab visits of the busiest CTA either way: every visit reloads
		   a slab and drains the pipeline */
		const int sms = svtgpu_sm_count();
		const int64_t per_cta = (m->nleaf + p.nchunks - 1) / p.nchunks;
		const int64_t by_leaves = ((per_cta + 511) / 512) * p.nstrips;
		const int sc = p.nstrips < sms ? p.nstrips : sms;
		const int64_t by_slabs = ((m->nleaf + 511) / 512) *
					 ((p.nstrips + sc - 1) / sc);
		const char *mode = svtgpu_env("SVTGPU_CP_MODE", "auto");
		/* (the partial results take nparts x nleaf x K doubles) */
		const bool fits = (double) sc * (double) m->nleaf * (double) K * 8.0
				  <= 8.0 * 1024 * 1024 * 1024;
		if (fits && (strcmp(mode, "slabs") == 0 ||
			     (strcmp(mode, "auto") == 0 && by_slabs * 3 < by_leaves * 2))) {
			p.slab_mode = 1;
			p.nchunks = sc;
			p.nparts = sc;
		}
	}
	p.ok = 1;
	return p;
}

/* d_rm: [nleaf][K] scratch, d_na: [nleaf] scratch */
template <typename T, bool LAC>
int launch_crossprod_strips(svtgpu_matrix *m, const CpPlan &p,
			    const double *Y, int64_t K, double *d_rm,
			    int32_t *d_na, cudaStream_t s)
{
	const int32_t *split = NULL;
	SVT_CHECK(svtgpu_ensure_split(m, p.nstrips, p.strip_rows, s, &split));
	if (!(p.panels && p.acc_tmem)) {
		SVT_CUDA(cudaMemsetAsync(d_rm, 0, 8 * (size_t) p.nparts *
					 (size_t) (m->nleaf * K), s));
		SVT_CUDA(cudaMemsetAsync(d_na, 0, 4 * (size_t) p.nparts *
					 (size_t) m->nleaf, s));
	}
	CpStripParams P;
	P.offs = m->d_offs;
	P.vals = LAC ? NULL : m->d_vals;
	P.leaf_ptr = m->d_leaf_ptr;
	P.split = split;
	P.nleaf = m->nleaf;
	P.nnz = m->nnz;
	P.nrow = m->nrow;
	P.nchunks = p.nchunks;
	P.nstrips = p.nstrips;
	P.strip_rows = p.strip_rows;
	P.K = (int) K;
	P.KP = p.KP;
	P.Y = Y;
	P.out = d_rm;
	P.leaf_na = d_na;
	P.slab_mode = p.slab_mode;
	if (p.panels) {
#define CPN_LAUNCH(NL, TM, BK) do { \
		SVT_CUDA(cudaFuncSetAttribute( \
			crossprod_panels<T, LAC, NL, TM, BK>, \
			cudaFuncAttributeMaxDynamicSharedMemorySize, \
			(int) p.smem)); \
		crossprod_panels<T, LAC, NL, TM, BK><<<(unsigned) p.nchunks, \
			p.warps * 32, p.smem, s>>>(P); \
	} while (0)
#define CPN_LAUNCH_NL(TM, BK) do { \
		if (p.KP > 50)      CPN_LAUNCH(4, TM, BK); \
		else if (p.KP > 48) CPN_LAUNCH(5, TM, BK); \
		else if (p.KP > 32) CPN_LAUNCH(3, TM, BK); \
		else if (p.KP > 16) CPN_LAUNCH(2, TM, BK); \
		else                CPN_LAUNCH(1, TM, BK); \
	} while (0)
		if (p.acc_tmem) {
			if (p.bulk) CPN_LAUNCH_NL(true, true);
			else        CPN_LAUNCH_NL(true, false);
		} else {
			if (p.bulk) CPN_LAUNCH_NL(false, true);
			else        CPN_LAUNCH_NL(false, false);
		}
#undef CPN_LAUNCH_NL
#undef CPN_LAUNCH
	} else if (p.KP > 32) {
		SVT_CUDA(cudaFuncSetAttribute(crossprod_strips<T, LAC, true>,
			cudaFuncAttributeMaxDynamicSharedMemorySize,
			(int) p.smem));
		crossprod_strips<T, LAC, true><<<(unsigned) p.nchunks,
			p.warps * 32, p.smem, s>>>(P);
	} else {
		SVT_CUDA(cudaFuncSetAttribute(crossprod_strips<T, LAC, false>,
			cudaFuncAttributeMaxDynamicSharedMemorySize,
			(int) p.smem));
		crossprod_strips<T, LAC, false><<<(unsigned) p.nchunks,
			p.warps * 32, p.smem, s>>>(P);
	}
	SVT_CUDA(cudaGetLastError());
	svtgpu_count_launch(1);
	return SVTGPU_OK;
}

int run_crossprod_strips(svtgpu_matrix *m, const CpPlan &p, const double *Y,
			 int64_t K, const SvtDenseColInfo *info, bool svt_left,
			 double *d_ans, cudaStream_t s)
{
	/* scratch: row-major sums, then the NA flags */
	void *scratch = NULL;
	const size_t rm_bytes = (8 * (size_t) p.nparts * (size_t) (m->nleaf * K) +
				 255) & ~(size_t) 255;
	SVT_CHECK(svtgpu_scratch(m, rm_bytes + 4 * (size_t) p.nparts *
				 (size_t) m->nleaf + 256, &scratch));
	double *d_rm = (double *) scratch;
	int32_t *d_na = (int32_t *) ((char *) scratch + rm_bytes);
	int rc;
	if (!(m->flags & SVTGPU_HAS_VALS))
		rc = launch_crossprod_strips<int32_t, true>(m, p, Y, K, d_rm,
							    d_na, s);
	else if (svt_is_double(m->val_type))
		rc = launch_crossprod_strips<double, false>(m, p, Y, K, d_rm,
							    d_na, s);
	else
		rc = launch_crossprod_strips<int32_t, false>(m, p, Y, K, d_rm,
							     d_na, s);
	SVT_CHECK(rc);
	crossprod_finish<<<grid_for(m->nleaf * K, 256), 256, 0, s>>>(d_rm, d_na,
		m->nleaf, K, p.nparts, svt_is_double(m->val_type), info, d_ans,
		svt_left ? 1 : 0);
	SVT_CUDA(cudaGetLastError());
	svtgpu_count_launch(1);
	return SVTGPU_OK;
}

template <typename T, bool LAC>
int launch_scatter(const svtgpu_matrix *m, const double *D, int64_t K,
		   bool check_nf, double *prod, int32_t *row_na,
		   unsigned long long *row_first, int32_t *hits,
		   cudaStream_t s)
{
	unsigned grid = grid_for(m->nleaf, 8);
	if (check_nf)
		matmul_scatter<T, LAC, true><<<grid, 256, 0, s>>>(m->d_offs,
			(const T *) m->d_vals, m->d_leaf_ptr, m->nleaf, D, K,
			prod, row_na, row_first, hits);
	else
		matmul_scatter<T, LAC, false><<<grid, 256, 0, s>>>(m->d_offs,
			(const T *) m->d_vals, m->d_leaf_ptr, m->nleaf, D, K,
			prod, row_na, row_first, hits);
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

/* Sparse second operand: leaves [k0, k0 + K) of `y` written as the row-major
 * dense block the product kernels take (zero-filled by the caller), with the
 * per-column NA / non-finite counts.  One warp per leaf. */
template <typename T, bool LACUNAR>
__global__ void __launch_bounds__(256)
expand_leaves_rowmajor(const int64_t *__restrict__ leaf_ptr,
		       const int32_t *__restrict__ offs,
		       const T *__restrict__ vals, int64_t k0, int64_t K,
		       double *__restrict__ out, SvtDenseColInfo *info)
{
	const int lane = threadIdx.x & 31;
	const int64_t warps = ((int64_t) gridDim.x * blockDim.x) >> 5;
	const int64_t gw = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	for (int64_t k = gw; k < K; k += warps) {
		const int64_t start = leaf_ptr[k0 + k], end = leaf_ptr[k0 + k + 1];
		int nf = 0, na = 0;
		for (int64_t e = start + lane; e < end; e += 32) {
			double d = 1.0;
			if (!LACUNAR) {
				const T v = vals[e];
				d = (double) v;
				if (sizeof(T) == 4) {
					const int bad = (int32_t) v == SVT_NA_INT;
					nf += bad; na += bad;
				} else {
					nf += !svt_isfinite(d);
					na += svt_is_na_real(d);
				}
			}
			out[(int64_t) offs[e] * K + k] = d;
		}
		nf = (int) svt_warp_sum((long long) nf);
		na = (int) svt_warp_sum((long long) na);
		if (lane == 0) {
			info[k].n_nonfinite = nf;
			info[k].n_na = na;
		}
	}
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
	const CpPlan plan = plan_crossprod_strips(m, K);
	if (plan.ok) {
		/* the scratch is about to be re-used: keep the (all-zero)
		   column info in its own small allocation */
		SvtDenseColInfo *d_info = NULL;
		SVT_CUDA(svt_malloc_async((void **) &d_info,
				sizeof(SvtDenseColInfo) * (size_t) K, s));
		cudaError_t e = cudaMemsetAsync(d_info, 0,
				sizeof(SvtDenseColInfo) * (size_t) K, s);
		int rc = e == cudaSuccess ? SVTGPU_OK
			: svtgpu_cuda_fail(e, "crossprod_dev memset", __FILE__,
					   __LINE__);
		if (rc == SVTGPU_OK)
			rc = run_crossprod_strips(m, plan, Y, K, d_info, true,
						  d_ans, s);
		svt_free_async(d_info, s);
		return rc;
	}
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
	SVT_CUDA(svt_malloc_async((void **) &d_buf, raw_bytes + rm_bytes +
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
	const CpPlan plan = plan_crossprod_strips(m, K);
	if (rc == SVTGPU_OK && plan.ok && !any_bad) {
		rc = run_crossprod_strips(m, plan, d_rm, K, d_info,
					  svt_on_left != 0, d_ans, s);
		int rc2 = svt_timer_end(&t, &m->tm.kernel_ms);
		if (rc == SVTGPU_OK)
			rc = rc2;
	} else if (rc == SVTGPU_OK) {
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
	svt_free_async(d_buf, s);
	return rc;
}

/* crossprod(x, y) with both operands sparse: ans = t(x) %*% y, nleaf(x) x
 * nleaf(y) column-major doubles on the host.  The reference pre-processes one
 * operand leaf by leaf into a dense column and runs the SVT x dense dot
 * products on it (crossprod2_{L,R}pp_*(), crossprod1_*(),
 * src/SparseMatrix_mult.c:560-929, :1037-1140) -- the results are those of
 * crossprod(x, as.matrix(y)) bit for bit (checked against the reference build
 * when the golden vectors are made).  Here blocks of <= 64 leaves of `y` are
 * expanded in HBM and go through the same product kernels. */
extern "C" int svtgpu_crossprod_svt(svtgpu_matrix *x, svtgpu_matrix *y,
				    double *ans)
{
	SVT_CHECK(svtgpu_require_device());
	SVT_ARG(x != NULL && y != NULL && ans != NULL,
		"svtgpu_crossprod_svt: NULL argument");
	SVT_ARG(x->nrow == y->nrow,
		"input SVT_SparseMatrix objects are non-conformable");
	SVT_ARG(x->val_type == y->val_type,
		"input SVT_SparseMatrix objects must have the same type() "
		"for now");
	SVT_CHECK(check_product_types(x, x->val_type, "crossprod"));
	SVT_CHECK(check_product_types(y, y->val_type, "crossprod"));
	SVT_CHECK(svtgpu_matrix_finish_upload(x));
	SVT_CHECK(svtgpu_matrix_finish_upload(y));
	x->tm.kernel_ms = x->tm.d2h_ms = 0.0;
	x->tm.d2h_bytes = 0.0;
	x->tm.launches = 0;
	const int64_t nx = x->nleaf, ny = y->nleaf, n = x->nrow;
	if (nx == 0 || ny == 0)
		return SVTGPU_OK;
	if (x->nnz == 0 && y->nnz == 0) {
		memset(ans, 0, sizeof(double) * (size_t) (nx * ny));
		return SVTGPU_OK;
	}
	/* Which operand becomes dense: the reference's rule
	   (src/SparseMatrix_mult.c:1075-1098, NULL SVTs :707-711,730-734) --
	   it decides NA vs NaN when one dot product meets both. */
	const bool dense_left = x != y && y->nnz > 0 &&
		(x->nnz == 0 || y->nnz * nx < x->nnz * ny);
	svtgpu_matrix *sp = dense_left ? y : x;    /* stays sparse */
	svtgpu_matrix *de = dense_left ? x : y;    /* expanded block by block */
	const int64_t nsp = sp->nleaf, nde = de->nleaf;
	cudaStream_t s = 0;
	const int64_t KB = 64;
	const size_t rm_bytes = (8 * (size_t) (n * KB) + 255) & ~(size_t) 255;
	const size_t info_bytes = (sizeof(SvtDenseColInfo) * (size_t) KB + 255) &
				  ~(size_t) 255;
	char *d_buf = NULL;
	SVT_CUDA(svt_malloc_async((void **) &d_buf, rm_bytes + info_bytes +
				 8 * (size_t) (nsp * KB), s));
	double *d_rm = (double *) d_buf;
	SvtDenseColInfo *d_info = (SvtDenseColInfo *) (d_buf + rm_bytes);
	double *d_ans = (double *) (d_buf + rm_bytes + info_bytes);
	const bool de_lac = !(de->flags & SVTGPU_HAS_VALS);
	const bool dbl = svt_is_double(de->val_type);
	const int64_t l0 = svtgpu_launch_count();
	int rc = SVTGPU_OK;
	double kernel_ms = 0.0;
	for (int64_t k0 = 0; k0 < nde && rc == SVTGPU_OK; k0 += KB) {
		const int64_t K = nde - k0 < KB ? nde - k0 : KB;
		cudaError_t e = cudaMemsetAsync(d_rm, 0, 8 * (size_t) (n * K), s);
		if (e == cudaSuccess)
			e = cudaMemsetAsync(d_info, 0,
					    sizeof(SvtDenseColInfo) * (size_t) K, s);
		SvtTimer t;
		bool timing = false;
		if (e == cudaSuccess) {
			if (svt_timer_begin(&t, s) == SVTGPU_OK)
				timing = true;
			else
				e = cudaErrorUnknown;
		}
		if (e == cudaSuccess && de->nnz > 0) {
			const unsigned grid = (unsigned) ((K + 7) / 8);
			if (de_lac)
				expand_leaves_rowmajor<int32_t, true><<<grid, 256, 0, s>>>(
					de->d_leaf_ptr, de->d_offs, NULL, k0, K,
					d_rm, d_info);
			else if (dbl)
				expand_leaves_rowmajor<double, false><<<grid, 256, 0, s>>>(
					de->d_leaf_ptr, de->d_offs,
					(const double *) de->d_vals, k0, K, d_rm,
					d_info);
			else
				expand_leaves_rowmajor<int32_t, false><<<grid, 256, 0, s>>>(
					de->d_leaf_ptr, de->d_offs,
					(const int32_t *) de->d_vals, k0, K, d_rm,
					d_info);
			e = cudaGetLastError();
			svtgpu_count_launch(1);
		}
		SvtDenseColInfo h[64];
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(h, d_info, sizeof(SvtDenseColInfo) *
					    (size_t) K, cudaMemcpyDeviceToHost, s);
		if (e == cudaSuccess)
			e = cudaStreamSynchronize(s);
		if (e != cudaSuccess) {
			if (timing) {
				double ms;
				svt_timer_end(&t, &ms);
			}
			rc = svtgpu_cuda_fail(e, "crossprod_svt expand", __FILE__,
					      __LINE__);
			break;
		}
		int any_bad = 0;
		for (int64_t k = 0; k < K; k++)
			if (h[k].n_nonfinite != 0)
				any_bad = 1;
		if (de->nnz == 0) {
			/* a NULL SVT: the reference runs its "fictive matrix of
			   zeros" routines (crossprod2_SVT_mat0_*(),
			   crossprod2_mat0_SVT_*(), src/SparseMatrix_mult.c:
			   558-629), whose NA rule differs from a dense column
			   of zeros: mark the columns */
			e = cudaMemsetAsync(d_info, 0xFF,
					    sizeof(SvtDenseColInfo) * (size_t) K, s);
			if (e != cudaSuccess) {
				double ms;
				svt_timer_end(&t, &ms);
				rc = svtgpu_cuda_fail(e, "crossprod_svt info",
						      __FILE__, __LINE__);
				break;
			}
		}
		const bool left = !dense_left;   /* the sparse side is x */
		/* (the sparse side always has nonzeros: an all-zero operand
		   is the dense one by the rule above) */
		const CpPlan plan = plan_crossprod_strips(sp, K);
		if (plan.ok && !any_bad)
			rc = run_crossprod_strips(sp, plan, d_rm, K, d_info, left,
						  d_ans, s);
		else if (!(sp->flags & SVTGPU_HAS_VALS))
			rc = launch_gather<int32_t, true>(sp, d_rm, K, d_info,
					any_bad, left, d_ans, s);
		else if (svt_is_double(sp->val_type))
			rc = launch_gather<double, false>(sp, d_rm, K, d_info,
					any_bad, left, d_ans, s);
		else
			rc = launch_gather<int32_t, false>(sp, d_rm, K, d_info,
					any_bad, left, d_ans, s);
		double ms = 0.0;
		int rc2 = svt_timer_end(&t, &ms);
		kernel_ms += ms;
		if (rc == SVTGPU_OK)
			rc = rc2;
		if (rc == SVTGPU_OK) {
			if (left)   /* nx x K block = columns [k0, k0 + K) */
				e = cudaMemcpyAsync(ans + k0 * nx, d_ans,
						    8 * (size_t) (nx * K),
						    cudaMemcpyDeviceToHost, s);
			else        /* K x ny block = rows [k0, k0 + K) */
				e = cudaMemcpy2DAsync(ans + k0, 8 * (size_t) nx,
						      d_ans, 8 * (size_t) K,
						      8 * (size_t) K, (size_t) ny,
						      cudaMemcpyDeviceToHost, s);
			if (e == cudaSuccess)
				e = cudaStreamSynchronize(s);
			if (e != cudaSuccess)
				rc = svtgpu_cuda_fail(e, "crossprod_svt D2H",
						      __FILE__, __LINE__);
		}
	}
	if (rc == SVTGPU_OK && x == y) {
		/* crossprod(x): the reference computes the pair (j, i), j < i,
		   once -- with leaf j pre-processed (dense when it is finite,
		   else sparse x sparse with "NA wins") -- and mirrors it
		   (compute_sym_dotprods_double(), src/SparseMatrix_mult.c:
		   826-852).  ans[i, j] (leaf i against dense column j) is that
		   value; ans[j, i] differs when leaf j holds a NaN before an
		   NA and column i is clean: mirror the lower triangle. */
		for (int64_t j = 0; j < nx; j++)
			for (int64_t i = j + 1; i < nx; i++)
				ans[j + i * nx] = ans[i + j * nx];
	}
	x->tm.kernel_ms = kernel_ms;
	x->tm.launches = (int) (svtgpu_launch_count() - l0);
	x->tm.d2h_bytes = 8.0 * (double) (nx * ny);
	svt_free_async(d_buf, s);
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
	/* svt %*% D = crossprod(t(svt), D): with the cached device transpose
	   the product is the shared-memory slab gather instead of 2 K fp64
	   reductions in L2 per nonzero */
	/* device-resident shard: the transpose is built once and reused */
	if (strcmp(svtgpu_env("SVTGPU_MM_IMPL", "auto"), "scatter") != 0) {
		svtgpu_matrix *tm = NULL;
		SVT_CHECK(svtgpu_ensure_transpose(m, s, &tm));
		if (tm != NULL) {
			const CpPlan plan = plan_crossprod_strips(tm, K);
			if (plan.ok) {
				SvtDenseColInfo *d_info = NULL;
				SVT_CUDA(svt_malloc_async((void **) &d_info,
					sizeof(SvtDenseColInfo) * (size_t) K, s));
				cudaError_t e = cudaMemsetAsync(d_info, 0,
					sizeof(SvtDenseColInfo) * (size_t) K, s);
				int rc = e == cudaSuccess ? SVTGPU_OK
					: svtgpu_cuda_fail(e, "matmul_dev memset",
							   __FILE__, __LINE__);
				if (rc == SVTGPU_OK)
					rc = run_crossprod_strips(tm, plan,
						(const double *) d_d_rowmajor, K,
						d_info, false, d_ans_rowmajor, s);
				svt_free_async(d_info, s);
				return rc;
			}
		}
	}
	/* device form assumes clean operands: NA flags are not tracked */
	void *scratch = NULL;
	SVT_CHECK(svtgpu_scratch(m, 4 * (size_t) m->nrow + 64, &scratch));
	int32_t *row_na = (int32_t *) scratch;
	const double *D = (const double *) d_d_rowmajor;
	if (!(m->flags & SVTGPU_HAS_VALS))
		return launch_scatter<int32_t, true>(m, D, K, false,
				d_ans_rowmajor, row_na, NULL, NULL, s);
	if (svt_is_double(m->val_type))
		return launch_scatter<double, false>(m, D, K, false,
				d_ans_rowmajor, row_na, NULL, NULL, s);
	return launch_scatter<int32_t, false>(m, D, K, false, d_ans_rowmajor,
					      row_na, NULL, NULL, s);
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
	const size_t first_bytes = (8 * (size_t) nrow + 255) & ~(size_t) 255;
	char *d_buf = NULL;
	SVT_CUDA(svt_malloc_async((void **) &d_buf, raw_bytes + rm_bytes +
				 info_bytes + 2 * prod_bytes + na_bytes +
				 first_bytes + 4 * nout + 256, s));
	char *p = d_buf;
	void *d_raw = p; p += raw_bytes;
	double *d_rm = (double *) p; p += rm_bytes;
	SvtDenseColInfo *d_info = (SvtDenseColInfo *) p; p += info_bytes;
	double *d_prod = (double *) p; p += prod_bytes;
	double *d_ans = (double *) p; p += prod_bytes;
	int32_t *d_row_na = (int32_t *) p; p += na_bytes;
	unsigned long long *d_row_first = (unsigned long long *) p;
	p += first_bytes;
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
	/* A stateless call pays for its own transpose: ~70 ms at 2.3e9
	   nonzeros (transpose_blocks), after which a dense column costs ~1 ms
	   in the slab kernel against ~3.5 ms as L2 reductions (measured, both
	   linear in nnz) -- worth it from about 30 columns on.  A handle that
	   already has its transpose always uses it. */
	bool done = false;
	const char *mm_impl = svtgpu_env("SVTGPU_MM_IMPL", "auto");
	const int64_t k_min = atoll(svtgpu_env("SVTGPU_MM_TRANSPOSE_MIN_K", "32"));
	if (rc == SVTGPU_OK && !any_bad && n > 0 &&
	    strcmp(mm_impl, "scatter") != 0 &&
	    (m->transposed != NULL || strcmp(mm_impl, "transpose") == 0 ||
	     (K >= k_min && K <= 64))) {
		svtgpu_matrix *tm = NULL;
		rc = svtgpu_ensure_transpose(m, s, &tm);
		if (rc == SVTGPU_OK && tm != NULL) {
			const CpPlan plan = plan_crossprod_strips(tm, K);
			if (plan.ok) {
				rc = run_crossprod_strips(tm, plan, d_rm, K,
						d_info, true, d_ans, s);
				done = true;
			}
		}
	}
	if (rc == SVTGPU_OK && !done) {
		cudaError_t e = cudaMemsetAsync(d_prod, 0, prod_bytes, s);
		if (e == cudaSuccess)
			e = cudaMemsetAsync(d_row_na, 0, na_bytes, s);
		if (e == cudaSuccess)
			e = cudaMemsetAsync(d_row_first, 0xFF, first_bytes, s);
		if (e == cudaSuccess && any_bad)
			e = cudaMemsetAsync(d_hits, 0, 4 * nout, s);
		if (e != cudaSuccess)
			rc = svtgpu_cuda_fail(e, "matmul memset", __FILE__,
					      __LINE__);
	}
	if (rc == SVTGPU_OK && !done && m->nnz > 0) {
		int32_t *hits = any_bad ? d_hits : NULL;
		if (!(m->flags & SVTGPU_HAS_VALS))
			rc = launch_scatter<int32_t, true>(m, d_rm, K, any_bad,
					d_prod, d_row_na, d_row_first, hits, s);
		else if (svt_is_double(m->val_type))
			rc = launch_scatter<double, false>(m, d_rm, K, any_bad,
					d_prod, d_row_na, d_row_first, hits, s);
		else
			rc = launch_scatter<int32_t, false>(m, d_rm, K,
					any_bad, d_prod, d_row_na, d_row_first,
					hits, s);
	}
	if (rc == SVTGPU_OK && !done) {
		matmul_finalize<<<grid_for((int64_t) nout, 256), 256, 0, s>>>(
			d_prod, d_row_na, d_row_first, any_bad ? d_hits : NULL,
			nrow, K,
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
	svt_free_async(d_buf, s);
	return rc;
}
