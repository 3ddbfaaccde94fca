/* rowsum() / colsum() of an SVT_SparseMatrix: sums of rows (columns) that
 * share a group label.  Replaces C_rowsum_SVT / C_colsum_SVT and the loops
 * behind them (reference src/rowsum_methods.c:44-125, :148-255, :281-326,
 * :364-409).
 *
 *   rowsum: out[g, j] = sum of x[i, j] over rows i with group[i] == g
 *           -> ngroup x ncol.  Leaves are independent: one warp per leaf
 *           with ngroup accumulators in shared memory, one coalesced write of
 *           the leaf's output column.
 *   colsum: out[i, g] = sum of x[i, j] over columns j with group[j] == g
 *           -> nrow x ngroup.  One warp per leaf scatters into the output
 *           column of the leaf's group with L2 reductions, a second kernel
 *           turns the accumulators into R's answer.
 *
 * Integer semantics.  The reference adds sequentially with an overflow check
 * after every addition (safe_int_add() for rowsum, the double-typed range
 * check of add_sparse_vec_to_ints() for colsum): NA is sticky, a partial sum
 * outside [-INT_MAX, INT_MAX] becomes NA and raises a warning.  The kernels
 * accumulate the exact sum S and A = sum |x| in 64 bits, plus the number of
 * NAs.  While A <= INT_MAX no partial sum can leave the range, whatever the
 * order, so the answer is S (or NA when an NA was met and na.rm is FALSE).
 * Cells with A > INT_MAX (possible only with values around 1e9) are replayed
 * sequentially by one thread with the reference's exact rules.
 *
 * Double semantics.  `out += v` in storage order; of several NA / NaN terms
 * the LAST one decides between NA_real_ and NaN (operand order of the
 * reference build, see svt_semantics.h), so the kernels keep the position of
 * the last NA and of the last NaN beside the sum of the regular values (which
 * yields the fresh NaN of Inf - Inf by itself).  The regular values are added
 * with atomics, i.e. in a different order than the reference: results agree
 * to rounding (<= 1e-12 relative), not bit for bit.
 */
#include <cuda_runtime.h>

#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/svtgpu.h"
#include "svtgpu_internal.h"
#include "svt_semantics.h"
#include "svt_ptx.cuh"

namespace {

struct GroupSumParams {
	const int32_t *offs;
	const void *vals;        /* NULL: lacunar (all ones) */
	const int64_t *leaf_ptr;
	const int32_t *group;    /* 0-based, NA already mapped to ngroup - 1 */
	int64_t nrow, nleaf;
	int ngroup;
	int narm;
	void *out;               /* int32 or double, column-major */
	/* colsum accumulators, nrow * ngroup each */
	long long *acc_sum;
	unsigned long long *acc_abs;
	int32_t *acc_a, *acc_b;  /* int: #NA, unused; double: last NA / NaN */
	int32_t *overflow;
	int vec_ok;              /* offs 16-byte aligned: 16-byte loads allowed */
};

/* one step of the reference's integer accumulation; returns the new cell */
__device__ __forceinline__ int rowsum_int_step(int cell, int v, int narm,
					       int *ovflow)
{
	/* compute_rowsum_ints(), src/rowsum_methods.c:66-84 */
	if (narm && v == SVT_NA_INT)
		return cell;
	if (cell == SVT_NA_INT || v == SVT_NA_INT)
		return SVT_NA_INT;
	if ((v > 0 && cell > INT_MAX - v) || (v < 0 && cell < -INT_MAX - v)) {
		*ovflow = 1;
		return SVT_NA_INT;
	}
	return cell + v;
}

__device__ __forceinline__ int colsum_int_step(int cell, int v, int narm,
					       int *ovflow)
{
	/* add_sparse_vec_to_ints(), src/rowsum_methods.c:172-199 */
	if (cell == SVT_NA_INT)
		return cell;
	if (v == SVT_NA_INT)
		return narm ? cell : SVT_NA_INT;
	const double y = (double) cell + (double) v;
	if (-(double) INT_MAX <= y && y <= (double) INT_MAX)
		return (int) y;
	*ovflow = 1;
	return SVT_NA_INT;
}

/* ---- rowsum: warp per leaf, accumulators in shared memory ---- */

template <bool LACUNAR>
__global__ void __launch_bounds__(256)
rowsum_int(GroupSumParams P)
{
	extern __shared__ __align__(16) unsigned char smem[];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int W = blockDim.x >> 5;
	const int G = P.ngroup;
	/* per warp: sum[G] (int64) | abs[G] (uint64) | nna[G] (int32) */
	unsigned char *base = smem + (size_t) warp * ((size_t) G * 20 + 12);
	base = (unsigned char *) (((uintptr_t) base + 7) & ~(uintptr_t) 7);
	unsigned long long *sum = (unsigned long long *) base;
	unsigned long long *abs_ = sum + G;
	int *nna = (int *) (abs_ + G);
	const int32_t *vals = (const int32_t *) P.vals;
	int32_t *out = (int32_t *) P.out;

	for (int64_t leaf = (int64_t) blockIdx.x * W + warp; leaf < P.nleaf;
	     leaf += (int64_t) gridDim.x * W) {
		const int64_t start = P.leaf_ptr[leaf];
		const int64_t end = P.leaf_ptr[leaf + 1];
		for (int g = lane; g < G; g += 32) {
			sum[g] = 0;
			abs_[g] = 0;
			nna[g] = 0;
		}
		__syncwarp();
		for (int64_t e = start + lane; e < end; e += 32) {
			const int g = P.group[P.offs[e]];
			const int x = LACUNAR ? 1 : vals[e];
			if (x == SVT_NA_INT) {
				if (!P.narm)
					atomicAdd(&nna[g], 1);
				continue;
			}
			atomicAdd(&sum[g], (unsigned long long) (long long) x);
			atomicAdd(&abs_[g], (unsigned long long)
					    (x < 0 ? -(long long) x : (long long) x));
		}
		__syncwarp();
		for (int g = lane; g < G; g += 32) {
			int r;
			if (abs_[g] <= (unsigned long long) INT_MAX) {
				r = nna[g] > 0 ? SVT_NA_INT
					       : (int) (long long) sum[g];
			} else {
				/* values around 1e9: the reference's order
				   decides */
				int ov = 0;
				r = 0;
				for (int64_t e = start; e < end; e++)
					if (P.group[P.offs[e]] == g)
						r = rowsum_int_step(r,
							LACUNAR ? 1 : vals[e],
							P.narm, &ov);
				if (ov)
					atomicOr(P.overflow, 1);
			}
			out[leaf * G + g] = r;
		}
		__syncwarp();
	}
}

template <bool LACUNAR>
__global__ void __launch_bounds__(256)
rowsum_double(GroupSumParams P)
{
	extern __shared__ __align__(16) unsigned char smem[];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int W = blockDim.x >> 5;
	const int G = P.ngroup;
	/* per warp: sum[G] (double) | last NA[G] | last NaN[G] (int32) */
	unsigned char *base = smem + (size_t) warp * ((size_t) G * 16 + 8);
	base = (unsigned char *) (((uintptr_t) base + 7) & ~(uintptr_t) 7);
	double *sum = (double *) base;
	int *last_na = (int *) (sum + G);
	int *last_nan = last_na + G;
	const double *vals = (const double *) P.vals;
	double *out = (double *) P.out;

	for (int64_t leaf = (int64_t) blockIdx.x * W + warp; leaf < P.nleaf;
	     leaf += (int64_t) gridDim.x * W) {
		const int64_t start = P.leaf_ptr[leaf];
		const int64_t end = P.leaf_ptr[leaf + 1];
		for (int g = lane; g < G; g += 32) {
			sum[g] = 0.0;
			last_na[g] = -1;
			last_nan[g] = -1;
		}
		__syncwarp();
		for (int64_t e = start + lane; e < end; e += 32) {
			const int g = P.group[P.offs[e]];
			const double x = LACUNAR ? 1.0 : vals[e];
			if (svt_isnan(x)) {
				if (!P.narm)
					atomicMax(svt_is_na_real(x) ? &last_na[g]
								    : &last_nan[g],
						  (int) (e - start));
				continue;
			}
			atomicAdd(&sum[g], x);
		}
		__syncwarp();
		for (int g = lane; g < G; g += 32) {
			double r = sum[g];
			if (last_na[g] >= 0 || last_nan[g] >= 0)
				r = last_na[g] > last_nan[g] ? svt_na_real()
							     : svt_nan();
			out[leaf * G + g] = r;
		}
		__syncwarp();
	}
}

/* ---- rowsum, few groups: lane-private accumulators, no atomics ----
 * Shared memory per warp: cell[g][lane] of 16 bytes (bank = lane: conflict
 * free).  Integer: {sum (int64), sum |x| (63 bits) | NA seen (bit 63)};
 * double: the sum, with the position of the last NA / NaN of each group (rare)
 * kept per warp by shared-memory atomicMax.  At the end of a leaf lane g folds
 * the 32 cells of group g, reading them skewed so that the lanes hit distinct
 * banks: a fixed order, deterministic. */
struct __align__(16) IntCell { long long sum; unsigned long long abs_na; };

template <bool LACUNAR>
__global__ void __launch_bounds__(256)
rowsum_int_private(GroupSumParams P)
{
	extern __shared__ __align__(16) unsigned char smem[];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int W = blockDim.x >> 5;
	const int G = P.ngroup;
	IntCell *cell = (IntCell *) smem + (size_t) warp * G * 32;
	const int32_t *vals = (const int32_t *) P.vals;
	int32_t *out = (int32_t *) P.out;
	const unsigned long long NA_BIT = 1ull << 63;

	for (int64_t leaf = (int64_t) blockIdx.x * W + warp; leaf < P.nleaf;
	     leaf += (int64_t) gridDim.x * W) {
		const int64_t start = P.leaf_ptr[leaf];
		const int64_t end = P.leaf_ptr[leaf + 1];
		for (int g = 0; g < G; g++) {
			cell[g * 32 + lane].sum = 0;
			cell[g * 32 + lane].abs_na = 0;
		}
#pragma unroll 4
		for (int64_t e = start + lane; e < end; e += 32) {
			const int g = P.group[P.offs[e]];
			const int x = LACUNAR ? 1 : vals[e];
			IntCell c = cell[g * 32 + lane];
			if (x == SVT_NA_INT) {
				if (P.narm)
					continue;
				c.abs_na |= NA_BIT;
			} else {
				c.sum += x;
				c.abs_na += (unsigned long long)
					(x < 0 ? -(long long) x : (long long) x);
			}
			cell[g * 32 + lane] = c;
		}
		__syncwarp();
		for (int g0 = 0; g0 < G; g0 += 32) {
			const int g = g0 + lane;
			if (g < G) {
				long long sum = 0;
				unsigned long long a = 0;
				bool na = false;
				for (int i = 0; i < 32; i++) {
					const IntCell c =
						cell[g * 32 + ((i + lane) & 31)];
					sum += c.sum;
					a += c.abs_na & ~NA_BIT;
					na |= (c.abs_na & NA_BIT) != 0;
				}
				int r;
				if (a <= (unsigned long long) INT_MAX) {
					r = na ? SVT_NA_INT : (int) sum;
				} else {
					int ov = 0;
					r = 0;
					for (int64_t e = start; e < end; e++)
						if (P.group[P.offs[e]] == g)
							r = rowsum_int_step(r,
							    LACUNAR ? 1 : vals[e],
							    P.narm, &ov);
					if (ov)
						atomicOr(P.overflow, 1);
				}
				out[leaf * G + g] = r;
			}
		}
		__syncwarp();
	}
}

template <bool LACUNAR, bool GROUPS_ON_CHIP>
__global__ void __launch_bounds__(256)
rowsum_double_private(GroupSumParams P)
{
	extern __shared__ __align__(16) unsigned char smem[];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int W = blockDim.x >> 5;
	const int G = P.ngroup;
	/* [W][G][32] sums | [W][2][G] last NA / NaN positions | group bytes */
	double *cell = (double *) smem + (size_t) warp * G * 32 + lane;
	int *last = (int *) ((double *) smem + (size_t) W * G * 32) +
		    (size_t) warp * 2 * G;
	unsigned char *grp8 = (unsigned char *)
		((int *) ((double *) smem + (size_t) W * G * 32) +
		 (size_t) W * 2 * G);
	if (GROUPS_ON_CHIP) {
		for (int64_t r = threadIdx.x; r < P.nrow; r += blockDim.x)
			grp8[r] = (unsigned char) P.group[r];
		__syncthreads();
	}
	const double *vals = (const double *) P.vals;
	double *out = (double *) P.out;
	constexpr int U = 8;

	for (int64_t leaf = (int64_t) blockIdx.x * W + warp; leaf < P.nleaf;
	     leaf += (int64_t) gridDim.x * W) {
		const int64_t start = P.leaf_ptr[leaf];
		const int64_t end = P.leaf_ptr[leaf + 1];
		for (int g = 0; g < G; g++)
			cell[g * 32] = 0.0;
		for (int g = lane; g < 2 * G; g += 32)
			last[g] = -1;
		__syncwarp();
		for (int64_t base = start + lane; base < end; base += 32 * U) {
			int o[U], g[U];
			double x[U];
#pragma unroll
			for (int k = 0; k < U; k++) {
				const int64_t e = base + k * 32;
				const bool ok = e < end;
				o[k] = ok ? P.offs[e] : 0;
				x[k] = ok ? (LACUNAR ? 1.0 : vals[e]) : 0.0;
			}
#pragma unroll
			for (int k = 0; k < U; k++)
				g[k] = GROUPS_ON_CHIP ? (int) grp8[o[k]]
						      : P.group[o[k]];
#pragma unroll
			for (int k = 0; k < U; k++) {
				if (svt_isnan(x[k])) {
					/* rare: remember where the last NA /
					   NaN of the group sits */
					if (!P.narm)
						atomicMax(&last[(svt_is_na_real(x[k])
								 ? 0 : G) + g[k]],
							  (int) (base + k * 32 - start));
					continue;
				}
				cell[g[k] * 32] += x[k];
			}
		}
		__syncwarp();
		for (int g0 = 0; g0 < G; g0 += 32) {
			const int gg = g0 + lane;
			if (gg < G) {
				const double *row = (const double *) smem +
					(size_t) warp * G * 32 + gg * 32;
				/* a fixed (skewed) order: deterministic */
				double sum = 0.0;
				for (int i = 0; i < 32; i++)
					sum += row[(i + lane) & 31];
				const int na = last[gg], nan = last[G + gg];
				if (na >= 0 || nan >= 0)
					sum = na > nan ? svt_na_real() : svt_nan();
				out[leaf * G + gg] = sum;
			}
		}
		__syncwarp();
	}
}

/* ---- rowsum, integer counts: <= 64 groups and nrow * max|x| <= INT_MAX
 * (no sum can leave the int range, known from the handle).  One int32 cell
 * per (group, lane), NA groups in a register bit mask, eight nonzeros per
 * lane in flight. */
template <bool LACUNAR, bool GROUPS_ON_CHIP>
__global__ void __launch_bounds__(256)
rowsum_int_small(GroupSumParams P)
{
	extern __shared__ __align__(16) unsigned char smem[];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int W = blockDim.x >> 5;
	const int G = P.ngroup;
	int *cell = (int *) smem + (size_t) warp * G * 32 + lane;
	/* the group of every row as one byte, behind the cells */
	unsigned char *grp8 = smem + (size_t) W * G * 32 * sizeof(int);
	if (GROUPS_ON_CHIP) {
		for (int64_t r = threadIdx.x; r < P.nrow; r += blockDim.x)
			grp8[r] = (unsigned char) P.group[r];
		__syncthreads();
	}
	const int32_t *vals = (const int32_t *) P.vals;
	int32_t *out = (int32_t *) P.out;
	constexpr int U = 8;

	for (int64_t leaf = (int64_t) blockIdx.x * W + warp; leaf < P.nleaf;
	     leaf += (int64_t) gridDim.x * W) {
		const int64_t start = P.leaf_ptr[leaf];
		const int64_t end = P.leaf_ptr[leaf + 1];
		for (int g = 0; g < G; g++)
			cell[g * 32] = 0;
		unsigned long long na_mask = 0;
		for (int64_t base = start + lane; base < end; base += 32 * U) {
			int o[U], x[U], g[U];
#pragma unroll
			for (int k = 0; k < U; k++) {
				const int64_t e = base + k * 32;
				const bool ok = e < end;
				o[k] = ok ? P.offs[e] : 0;
				x[k] = ok ? (LACUNAR ? 1 : vals[e]) : 0;
			}
#pragma unroll
			for (int k = 0; k < U; k++)
				g[k] = GROUPS_ON_CHIP ? (int) grp8[o[k]]
						      : P.group[o[k]];
#pragma unroll
			for (int k = 0; k < U; k++) {
				if (x[k] == SVT_NA_INT) {
					if (!P.narm)
						na_mask |= 1ull << g[k];
					x[k] = 0;
				}
				cell[g[k] * 32] += x[k];
			}
		}
		__syncwarp();
		const unsigned int na_lo = __reduce_or_sync(SVT_FULL_MASK,
						(unsigned int) na_mask);
		const unsigned int na_hi = __reduce_or_sync(SVT_FULL_MASK,
						(unsigned int) (na_mask >> 32));
		const unsigned long long na_all =
			((unsigned long long) na_hi << 32) | na_lo;
		for (int g0 = 0; g0 < G; g0 += 32) {
			const int gg = g0 + lane;
			if (gg < G) {
				const int *row = (const int *) smem +
						 (size_t) warp * G * 32 + gg * 32;
				int sum = 0;
				for (int i = 0; i < 32; i++)
					sum += row[(i + lane) & 31];
				out[leaf * G + gg] = ((na_all >> gg) & 1)
						     ? SVT_NA_INT : sum;
			}
		}
		__syncwarp();
	}
}

/* ---- rowsum of a lacunar matrix in <= 16 groups: the leaf's group counts
 * live in REGISTERS.  A nonzero is only a row offset; its group comes from a
 * byte table in shared memory (one table per CTA of 32 warps, so 100,000 rows
 * leave room for full occupancy of the load pipeline: rowsum_int_small holds
 * 2 x 8 warps per SM there) and bumps a 4-bit field of one 64-bit register:
 * a byte load, a 64-bit shift and a 64-bit add per nonzero, no shared-memory
 * read-modify-write, no branch (entries past the end of a leaf point at a
 * table entry whose shift count is 64: the increment is zero).  After every
 * batch of 8 entries per lane the 4-bit fields are added to 8-bit fields of
 * two more registers, and those go to per-lane 32-bit totals (lane k of the
 * fold owns group k) at the end of the leaf or every 248 entries per lane.
 * (ncu on the first version -- branchy, 32-bit fields selected by four
 * compares: 44 warp instructions per 32 nonzeros, 88 % of the issue slots.) */
__device__ __forceinline__ uint64_t shl64_clamped(uint64_t x, uint32_t n)
{
	uint64_t r;
	asm("shl.b64 %0, %1, %2;" : "=l"(r) : "l"(x), "r"(n));
	return r;
}

__global__ void __launch_bounds__(1024, 1)
rowsum_lacunar_packed(GroupSumParams P)
{
	extern __shared__ __align__(16) unsigned char smem[];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int W = blockDim.x >> 5;
	const int G = P.ngroup;
	/* entry: bit position of the group's 4-bit field; entry nrow: 64 */
	unsigned char *tab = smem;
	for (int64_t r = threadIdx.x; r <= P.nrow; r += blockDim.x)
		tab[r] = r < P.nrow ? (unsigned char) (P.group[r] << 2) : 64;
	__syncthreads();
	const uint32_t tab_s = (uint32_t) __cvta_generic_to_shared(tab);
	const int none = (int) P.nrow;
	int32_t *out = (int32_t *) P.out;
	constexpr int U = 8;
	const int64_t stride = (int64_t) gridDim.x * W;

	/* The warp's leaves form one stream of batches of 32 x U offsets; the
	   loads of a batch are issued before the previous batch is counted, and
	   the bounds of a leaf one leaf before its first batch, so that the
	   warp always has loads in flight. */
	auto bounds = [&](int64_t lf, int64_t &b0, int64_t &b1) {
		b0 = b1 = 0;
		if (lf < P.nleaf) {
			b0 = P.leaf_ptr[lf];
			b1 = P.leaf_ptr[lf + 1];
		}
	};
	auto load = [&](int (&o)[U], int64_t base, int64_t end) {
		const int32_t *p = P.offs + base + lane;
		const int64_t left = end - base - lane;
		const int rem = left > 32 * U ? 32 * U : (int) left;
#pragma unroll
		for (int k = 0; k < U; k++)
			o[k] = k * 32 < rem ? p[k * 32] : none;
	};
	int64_t nx_leaf = (int64_t) blockIdx.x * W + warp;
	int64_t fb, fe, ns, ne;
	bounds(nx_leaf, fb, fe);
	int64_t ahead = nx_leaf + stride;
	bounds(ahead, ns, ne);
	int o_nx[U];
	if (nx_leaf < P.nleaf)
		load(o_nx, fb, fe);
	bool nx_last = fb + 32 * U >= fe;

	uint64_t ev = 0, od = 0;  /* 8-bit fields: groups 0, 2, ... | 1, 3, ... */
	int total = 0;            /* lane k: count of group k so far */
	int since = 0;
	auto spill = [&]() {
#pragma unroll
		for (int k = 0; k < 16; k++) {
			if (k >= G)
				break;
			const uint64_t c = (k & 1) ? od : ev;
			const int n = __reduce_add_sync(SVT_FULL_MASK,
					(int) ((c >> ((k >> 1) * 8)) & 0xFFu));
			if (lane == k)
				total += n;
		}
		ev = od = 0;
		since = 0;
	};
	while (nx_leaf < P.nleaf) {
		int o[U];
#pragma unroll
		for (int k = 0; k < U; k++)
			o[k] = o_nx[k];
		const int64_t leaf = nx_leaf;
		const bool last = nx_last;
		/* the next batch */
		fb += 32 * U;
		if (last) {
			nx_leaf = ahead;
			fb = ns;
			fe = ne;
			ahead += stride;
			bounds(ahead, ns, ne);
		}
		if (nx_leaf < P.nleaf)
			load(o_nx, fb, fe);
		nx_last = fb + 32 * U >= fe;
		/* count this one */
		if (since + U > 255)
			spill();
		since += U;
		uint64_t acc4 = 0;
#pragma unroll
		for (int k = 0; k < U; k++) {
			uint32_t t;
			asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t)
				     : "r"(tab_s + (uint32_t) o[k]));
			acc4 += shl64_clamped(1ull, t);
		}
		ev += acc4 & 0x0F0F0F0F0F0F0F0Full;
		od += (acc4 >> 4) & 0x0F0F0F0F0F0F0F0Full;
		if (last) {
			spill();
			if (lane < G)
				out[leaf * G + lane] = total;
			total = 0;
		}
	}
}

/* ---- colsum: warp per leaf, L2 reductions into the group's column ---- */

template <typename T, bool LACUNAR>
__global__ void __launch_bounds__(256)
colsum_scatter(GroupSumParams P)
{
	const int lane = threadIdx.x & 31;
	const int64_t warps = ((int64_t) gridDim.x * blockDim.x) >> 5;
	const int64_t gw = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const T *vals = (const T *) P.vals;
	for (int64_t leaf = gw; leaf < P.nleaf; leaf += warps) {
		const int64_t start = P.leaf_ptr[leaf];
		const int64_t end = P.leaf_ptr[leaf + 1];
		const int64_t col = (int64_t) P.group[leaf] * P.nrow;
		for (int64_t e = start + lane; e < end; e += 32) {
			const int64_t at = col + P.offs[e];
			if (sizeof(T) == 4) {
				const int x = LACUNAR ? 1 : (int) vals[e];
				if (x == SVT_NA_INT) {
					if (!P.narm)
						atomicAdd(&P.acc_a[at], 1);
					continue;
				}
				atomicAdd((unsigned long long *) &P.acc_sum[at],
					  (unsigned long long) (long long) x);
				atomicAdd(&P.acc_abs[at], (unsigned long long)
					  (x < 0 ? -(long long) x : (long long) x));
			} else {
				const double x = LACUNAR ? 1.0 : (double) vals[e];
				if (svt_isnan(x)) {
					if (!P.narm)
						atomicMax(svt_is_na_real(x)
							  ? &P.acc_a[at]
							  : &P.acc_b[at],
							  (int) leaf);
					continue;
				}
				atomicAdd(&((double *) P.out)[at], x);
			}
		}
	}
}

template <bool LACUNAR>
__global__ void __launch_bounds__(256)
colsum_finish_int(GroupSumParams P)
{
	const int64_t n = P.nrow * P.ngroup;
	const int32_t *vals = (const int32_t *) P.vals;
	int32_t *out = (int32_t *) P.out;
	for (int64_t at = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	     at < n; at += (int64_t) gridDim.x * blockDim.x) {
		if (P.acc_abs[at] <= (unsigned long long) INT_MAX) {
			out[at] = P.acc_a[at] > 0 ? SVT_NA_INT
						  : (int) P.acc_sum[at];
			continue;
		}
		/* values around 1e9: replay this cell in column order */
		const int g = (int) (at / P.nrow);
		const int row = (int) (at - (int64_t) g * P.nrow);
		int r = 0, ov = 0;
		for (int64_t leaf = 0; leaf < P.nleaf; leaf++) {
			if (P.group[leaf] != g)
				continue;
			int64_t lo = P.leaf_ptr[leaf], hi = P.leaf_ptr[leaf + 1];
			while (lo < hi) {
				const int64_t mid = lo + ((hi - lo) >> 1);
				if (P.offs[mid] < row) lo = mid + 1;
				else                   hi = mid;
			}
			if (lo < P.leaf_ptr[leaf + 1] && P.offs[lo] == row)
				r = colsum_int_step(r, LACUNAR ? 1 : vals[lo],
						    P.narm, &ov);
		}
		if (ov)
			atomicOr(P.overflow, 1);
		out[at] = r;
	}
}

__global__ void __launch_bounds__(256)
colsum_finish_double(GroupSumParams P)
{
	const int64_t n = P.nrow * P.ngroup;
	double *out = (double *) P.out;
	for (int64_t at = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	     at < n; at += (int64_t) gridDim.x * blockDim.x) {
		const int na = P.acc_a[at], nan = P.acc_b[at];
		if (na >= 0 || nan >= 0)
			out[at] = na > nan ? svt_na_real() : svt_nan();
	}
}

/* integer colsum when no cell can leave the int range (max |x| * nleaf <=
   INT_MAX, known from the handle): one 32-bit reduction per nonzero straight
   into the result, NA counted beside it */
template <bool LACUNAR>
__global__ void __launch_bounds__(256)
colsum_scatter_small(GroupSumParams P)
{
	const int lane = threadIdx.x & 31;
	const int64_t warps = ((int64_t) gridDim.x * blockDim.x) >> 5;
	const int64_t gw = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int32_t *vals = (const int32_t *) P.vals;
	int32_t *out = (int32_t *) P.out;
	for (int64_t leaf = gw; leaf < P.nleaf; leaf += warps) {
		const int64_t start = P.leaf_ptr[leaf];
		const int64_t end = P.leaf_ptr[leaf + 1];
		const int64_t col = (int64_t) P.group[leaf] * P.nrow;
#pragma unroll 4
		for (int64_t e = start + lane; e < end; e += 32) {
			const int64_t at = col + P.offs[e];
			const int x = LACUNAR ? 1 : vals[e];
			if (x == SVT_NA_INT) {
				if (!P.narm)
					atomicAdd(&P.acc_a[at], 1);
				continue;
			}
			atomicAdd(&out[at], x);
		}
	}
}

__global__ void __launch_bounds__(256)
colsum_finish_small(GroupSumParams P)
{
	const int64_t n = P.nrow * P.ngroup;
	int32_t *out = (int32_t *) P.out;
	for (int64_t at = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	     at < n; at += (int64_t) gridDim.x * blockDim.x)
		if (P.acc_a[at] > 0)
			out[at] = SVT_NA_INT;
}

/* integer colsum, bounded counts, nrow * 4 bytes fit one SM: the leaves are
   visited group by group (perm = leaves sorted by group, cut into pieces of
   one group each); a CTA sums a piece into nrow int32 cells of shared memory
   (rows are distinct inside a leaf, so only different warps can meet on a
   cell: shared-memory atomics) and adds its non-zero cells to the result
   once per piece. */
struct ColsumPiece { int32_t g, begin, end, pad; };

/* HALF16: non-negative counts whose sum over one piece fits 16 bits -- two
   rows share a 32-bit cell, so 100,000 rows fit one SM */
template <bool LACUNAR, bool HALF16>
__global__ void __launch_bounds__(1024, 1)
colsum_pieces_small(GroupSumParams P, const int32_t *__restrict__ perm,
		    const ColsumPiece *__restrict__ pieces, int npieces)
{
	extern __shared__ __align__(16) unsigned char smem[];
	int *acc = (int *) smem;
	const uint32_t acc_s = (uint32_t) __cvta_generic_to_shared(acc);
	const int64_t ncell = HALF16 ? (P.nrow + 1) / 2 : P.nrow;
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int W = blockDim.x >> 5;
	const int32_t *vals = (const int32_t *) P.vals;
	int32_t *out = (int32_t *) P.out;
	constexpr int U = 8;
	for (int p = blockIdx.x; p < npieces; p += gridDim.x) {
		const ColsumPiece pc = pieces[p];
		for (int64_t r = threadIdx.x; r < ncell; r += blockDim.x)
			acc[r] = 0;
		__syncthreads();
		const int64_t col = (int64_t) pc.g * P.nrow;
		/* the warp's leaves as one stream of batches of 32 x U entries:
		   the loads of a batch are issued before the previous one is
		   added, the bounds of a leaf one leaf ahead (leaf by leaf the
		   warp waited for perm -> leaf_ptr -> offsets at every leaf) */
		auto bounds = [&](int i, int64_t &b0, int64_t &b1) {
			b0 = b1 = 0;
			if (i < pc.end) {
				const int64_t leaf = perm[i];
				b0 = P.leaf_ptr[leaf];
				b1 = P.leaf_ptr[leaf + 1];
			}
		};
		auto load = [&](int (&o)[U], int (&x)[U], int64_t base,
				int64_t end) {
			const int32_t *po = P.offs + base + lane;
			const int32_t *pv = LACUNAR ? NULL : vals + base + lane;
			const int64_t left = end - base - lane;
			const int rem = left > 32 * U ? 32 * U : (int) left;
#pragma unroll
			for (int k = 0; k < U; k++) {
				const bool ok = k * 32 < rem;
				o[k] = ok ? po[k * 32] : 0;
				x[k] = ok ? (LACUNAR ? 1 : pv[k * 32]) : 0;
			}
		};
		int nx_i = pc.begin + warp;
		int64_t fb, fe, ns, ne;
		bounds(nx_i, fb, fe);
		int ahead = nx_i + W;
		bounds(ahead, ns, ne);
		if (LACUNAR && P.vec_ok) {
			/* lacunar leaves: offsets only, one atomic per 4 bytes --
			   the kernel waits on the MIO queue, which the load
			   instructions fill (see row_hist): 16-byte loads over
			   the aligned middle of a leaf, the < 4 entries at either
			   end by single lanes */
			auto bump = [&](uint32_t ou) {
				if (HALF16)
					asm volatile("red.shared.add.u32 [%0], %1;"
					    :: "r"(acc_s + ((ou << 1) & ~3u)),
					       "r"(1u << ((ou & 1u) << 4)) : "memory");
				else
					asm volatile("red.shared.add.u32 [%0], %1;"
					    :: "r"(acc_s + (ou << 2)), "r"(1u)
					    : "memory");
			};
			while (nx_i < pc.end) {
				const int64_t start = fb, end = fe;
				nx_i = ahead;
				fb = ns;
				fe = ne;
				ahead += W;
				bounds(ahead, ns, ne);
				const int64_t a4 = (start + 3) & ~(int64_t) 3;
				const int64_t b4 = end & ~(int64_t) 3;
				if (a4 >= b4) {
					for (int64_t e = start + lane; e < end; e += 32)
						bump((uint32_t) P.offs[e]);
					continue;
				}
				int oh = -1, ot = -1;
				if (start + lane < a4)
					oh = P.offs[start + lane];
				if (b4 + lane < end)
					ot = P.offs[b4 + lane];
				const int4 *q = (const int4 *) (P.offs + a4);
				const int nv = (int) ((b4 - a4) >> 2);
				for (int iv = lane; iv < nv; iv += 128) {
					int4 v[4];
#pragma unroll
					for (int k = 0; k < 4; k++)
						if (iv + 32 * k < nv)
							v[k] = q[iv + 32 * k];
#pragma unroll
					for (int k = 0; k < 4; k++) {
						if (iv + 32 * k < nv) {
							bump((uint32_t) v[k].x);
							bump((uint32_t) v[k].y);
							bump((uint32_t) v[k].z);
							bump((uint32_t) v[k].w);
						}
					}
				}
				if (oh >= 0) bump((uint32_t) oh);
				if (ot >= 0) bump((uint32_t) ot);
			}
		}
		int o_nx[U], x_nx[U];
		if (nx_i < pc.end)
			load(o_nx, x_nx, fb, fe);
		while (nx_i < pc.end) {
			int o[U], x[U];
#pragma unroll
			for (int k = 0; k < U; k++) {
				o[k] = o_nx[k];
				x[k] = x_nx[k];
			}
			fb += 32 * U;
			if (fb >= fe) {
				nx_i = ahead;
				fb = ns;
				fe = ne;
				ahead += W;
				bounds(ahead, ns, ne);
			}
			if (nx_i < pc.end)
				load(o_nx, x_nx, fb, fe);
			/* entries past the end of the leaf add 0 to cell 0 */
#pragma unroll
			for (int k = 0; k < U; k++) {
				if (!LACUNAR && x[k] == SVT_NA_INT) {
					if (!P.narm)
						atomicAdd(&P.acc_a[col + o[k]], 1);
					continue;
				}
				const uint32_t ou = (uint32_t) o[k];
				if (HALF16)
					asm volatile("red.shared.add.u32 [%0], %1;"
					    :: "r"(acc_s + ((ou << 1) & ~3u)),
					       "r"((uint32_t) x[k] << ((ou & 1u) << 4))
					    : "memory");
				else
					asm volatile("red.shared.add.u32 [%0], %1;"
					    :: "r"(acc_s + (ou << 2)),
					       "r"((uint32_t) x[k]) : "memory");
			}
		}
		__syncthreads();
		for (int64_t r = threadIdx.x; r < P.nrow; r += blockDim.x) {
			const int v = HALF16
				? (int) (((unsigned int) acc[r >> 1] >>
					  ((r & 1) * 16)) & 0xFFFFu)
				: acc[r];
			if (v != 0)
				atomicAdd(&out[col + r], v);
		}
		__syncthreads();
	}
}

/* group labels: 1-based with NA -> 0-based with NA = the last group
   (src/rowsum_methods.c:48-51).  Returns NULL + error on a bad label. */
int32_t *normalise_groups(const int32_t *group, int64_t n, int ngroup)
{
	int32_t *g = (int32_t *) malloc(sizeof(int32_t) * (size_t) (n > 0 ? n : 1));
	if (g == NULL) {
		svtgpu_set_error("out of host memory");
		return NULL;
	}
	for (int64_t i = 0; i < n; i++) {
		int32_t v = group[i];
		if (v == SVT_NA_INT) {
			if (ngroup < 1) {
				svtgpu_set_error("'ngroup' must be >= 1 when "
					"'group' contains missing values");
				free(g);
				return NULL;
			}
			v = ngroup;
		} else if (v < 1 || v > ngroup) {
			svtgpu_set_error("all non-NA values in 'group' must "
					 "be >= 1 and <= 'ngroup'");
			free(g);
			return NULL;
		}
		g[i] = v - 1;
	}
	return g;
}

int check_groupsum_input(const svtgpu_matrix *m, const void *group,
			 const void *out, const char *what)
{
	SVT_ARG(m != NULL && out != NULL, "%s: NULL argument", what);
	SVT_ARG(m->val_type == SVTGPU_INT || m->val_type == SVTGPU_DOUBLE,
		"rowsum() and colsum() do not support SVT_SparseMatrix "
		"objects of this type at the moment");
	SVT_ARG(group != NULL || m->nrow == 0 || m->nleaf == 0,
		"%s: NULL group", what);
	return SVTGPU_OK;
}

}  /* namespace */

extern "C" int svtgpu_rowsum(svtgpu_matrix *m, const int32_t *group,
			     int ngroup, int narm, void *out, int *overflow)
{
	SVT_CHECK(svtgpu_require_device());
	SVT_CHECK(check_groupsum_input(m, group, out, "svtgpu_rowsum"));
	SVT_ARG(ngroup >= 0, "svtgpu_rowsum: negative 'ngroup'");
	if (overflow != NULL)
		*overflow = 0;
	m->tm.kernel_ms = m->tm.d2h_ms = 0.0;
	m->tm.d2h_bytes = 0.0;
	m->tm.launches = 0;
	const int dbl = svt_is_double(m->val_type);
	const size_t esz = dbl ? 8 : 4;
	const size_t nout = (size_t) ngroup * (size_t) m->nleaf;
	int32_t *g0 = normalise_groups(group, m->nrow, ngroup);
	if (g0 == NULL)
		return SVTGPU_ERR_ARG;
	if (nout == 0) {
		free(g0);
		return SVTGPU_OK;
	}
	{   /* (not SVT_CHECK: g0 must not leak when the upload failed) */
		const int rcu = svtgpu_matrix_finish_upload(m);
		if (rcu != SVTGPU_OK) {
			free(g0);
			return rcu;
		}
	}
	if (m->nnz == 0) {
		free(g0);
		memset(out, 0, esz * nout);
		return SVTGPU_OK;
	}
	/* few groups: lane-private accumulators (16 B x 32 lanes per group and
	   warp) when at least 4 warps fit 200 KB; else shared-memory atomics on
	   one accumulator per group and warp */
	const char *impl = svtgpu_env("SVTGPU_ROWSUM_IMPL", "auto");
	cudaStream_t s = 0;
	/* integer counts in few groups: 32-bit lane-private cells */
	bool small = false;
	if (!dbl && ngroup <= 64 && strcmp(impl, "auto") == 0) {
		int rcb = svtgpu_ensure_absmax(m, s);
		if (rcb != SVTGPU_OK) {
			free(g0);
			return rcb;
		}
		const int64_t B = svtgpu_value_bound(m);
		small = B >= 0 && (B == 0 || m->nrow <= (int64_t) INT_MAX / B);
	}
	/* lacunar input in few groups: counts in packed registers */
	const bool packed = small && !(m->flags & SVTGPU_HAS_VALS) &&
		ngroup <= 16 && (size_t) m->nrow + 64 <= (size_t) 224 * 1024 &&
		m->nrow / 32 + 1 < (int64_t) INT32_MAX / 2 &&
		strcmp(svtgpu_env("SVTGPU_ROWSUM_LACUNAR", "packed"), "packed") == 0;
	const size_t priv_warp = small ? (size_t) ngroup * 32 * 4
			       : dbl ? (size_t) ngroup * (32 * 8 + 8)
				     : (size_t) ngroup * 32 * 16;
	const bool priv = small || (strcmp(impl, "atomic") != 0 &&
				    priv_warp * 4 <= (size_t) (200 * 1024));
	const size_t per_warp = priv ? priv_warp
				     : (size_t) ngroup * (dbl ? 16 : 20) + 16;
	int W = (int) ((size_t) (200 * 1024) / per_warp);
	if (W > 8) W = 8;
	if (W < 1) {
		free(g0);
		svtgpu_set_error("rowsum(): %d groups are more than the GPU "
				 "path holds on chip", ngroup);
		return SVTGPU_ERR_UNSUPPORTED;
	}
	/* the row -> group table rides along as bytes when it fits */
	const bool g_on_chip = (small || (priv && dbl)) && ngroup <= 255 &&
		per_warp * (size_t) W + (size_t) m->nrow + 64 <= (size_t) 200 * 1024;
	const size_t smem = per_warp * (size_t) W + 16 +
			    (g_on_chip ? (size_t) m->nrow + 48 : 0);
	
	void *scratch = NULL;
	const size_t g_bytes = (sizeof(int32_t) * (size_t) m->nrow + 255) &
			       ~(size_t) 255;
	int rc = svtgpu_scratch(m, g_bytes + 256 + esz * nout, &scratch);
	if (rc != SVTGPU_OK) {
		free(g0);
		return rc;
	}
	int32_t *d_group = (int32_t *) scratch;
	int32_t *d_ov = (int32_t *) ((char *) scratch + g_bytes);
	void *d_out = (char *) scratch + g_bytes + 256;
	cudaError_t e = cudaMemcpyAsync(d_group, g0, sizeof(int32_t) *
					(size_t) m->nrow, cudaMemcpyHostToDevice, s);
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(s);   /* g0 is pageable */
	free(g0);
	SVT_CUDA(e);
	SVT_CUDA(cudaMemsetAsync(d_ov, 0, 16, s));

	GroupSumParams P;
	memset(&P, 0, sizeof(P));
	P.offs = m->d_offs;
	P.vals = (m->flags & SVTGPU_HAS_VALS) ? m->d_vals : NULL;
	P.leaf_ptr = m->d_leaf_ptr;
	P.group = d_group;
	P.nrow = m->nrow;
	P.nleaf = m->nleaf;
	P.ngroup = ngroup;
	P.narm = narm != 0;
	P.out = d_out;
	P.overflow = d_ov;
	P.vec_ok = (((uintptr_t) m->d_offs) & 15) == 0 &&
		   strcmp(svtgpu_env("SVTGPU_COLSUM_VEC", "1"), "1") == 0;
	int64_t blocks = (m->nleaf + W - 1) / W;
	const int64_t cap = (int64_t) svtgpu_sm_count() * 8;
	if (blocks > cap) blocks = cap;
	const bool lac = P.vals == NULL;
	SvtTimer t;
	SVT_CHECK(svt_timer_begin(&t, s));
#define ROWSUM_LAUNCH(K) do { \
		SVT_CUDA(cudaFuncSetAttribute(K, \
			cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); \
		K<<<(unsigned) blocks, W * 32, smem, s>>>(P); \
	} while (0)
	if (packed) {
		const size_t psm = ((size_t) m->nrow + 64) & ~(size_t) 15;
		int64_t pb = (m->nleaf + 31) / 32;
		if (pb > svtgpu_sm_count()) pb = svtgpu_sm_count();
		SVT_CUDA(cudaFuncSetAttribute(rowsum_lacunar_packed,
			cudaFuncAttributeMaxDynamicSharedMemorySize, (int) psm));
		rowsum_lacunar_packed<<<(unsigned) pb, 1024, psm, s>>>(P);
	} else if (small && g_on_chip) {
		if (lac) ROWSUM_LAUNCH((rowsum_int_small<true, true>));
		else     ROWSUM_LAUNCH((rowsum_int_small<false, true>));
	} else if (small) {
		if (lac) ROWSUM_LAUNCH((rowsum_int_small<true, false>));
		else     ROWSUM_LAUNCH((rowsum_int_small<false, false>));
	} else if (priv && dbl && g_on_chip) {
		if (lac) ROWSUM_LAUNCH((rowsum_double_private<true, true>));
		else     ROWSUM_LAUNCH((rowsum_double_private<false, true>));
	} else if (priv && dbl) {
		if (lac) ROWSUM_LAUNCH((rowsum_double_private<true, false>));
		else     ROWSUM_LAUNCH((rowsum_double_private<false, false>));
	} else if (priv) {
		if (lac) ROWSUM_LAUNCH(rowsum_int_private<true>);
		else     ROWSUM_LAUNCH(rowsum_int_private<false>);
	} else if (dbl) {
		if (lac) ROWSUM_LAUNCH(rowsum_double<true>);
		else     ROWSUM_LAUNCH(rowsum_double<false>);
	} else {
		if (lac) ROWSUM_LAUNCH(rowsum_int<true>);
		else     ROWSUM_LAUNCH(rowsum_int<false>);
	}
#undef ROWSUM_LAUNCH
	cudaError_t le = cudaGetLastError();
	int rc2 = svt_timer_end(&t, &m->tm.kernel_ms);
	SVT_CUDA(le);
	SVT_CHECK(rc2);
	svtgpu_count_launch(1);
	m->tm.launches = 1;
	SVT_CHECK(svt_timer_begin(&t, s));
	int32_t h_ov = 0;
	SVT_CUDA(cudaMemcpyAsync(out, d_out, esz * nout, cudaMemcpyDeviceToHost, s));
	SVT_CUDA(cudaMemcpyAsync(&h_ov, d_ov, sizeof(int32_t),
				 cudaMemcpyDeviceToHost, s));
	SVT_CHECK(svt_timer_end(&t, &m->tm.d2h_ms));
	m->tm.d2h_bytes = (double) (esz * nout);
	if (overflow != NULL)
		*overflow = h_ov != 0;
	return SVTGPU_OK;
}

extern "C" int svtgpu_colsum(svtgpu_matrix *m, const int32_t *group,
			     int ngroup, int narm, void *out, int *overflow)
{
	SVT_CHECK(svtgpu_require_device());
	SVT_CHECK(check_groupsum_input(m, group, out, "svtgpu_colsum"));
	SVT_ARG(ngroup >= 0, "svtgpu_colsum: negative 'ngroup'");
	if (overflow != NULL)
		*overflow = 0;
	m->tm.kernel_ms = m->tm.d2h_ms = 0.0;
	m->tm.d2h_bytes = 0.0;
	m->tm.launches = 0;
	const int dbl = svt_is_double(m->val_type);
	const size_t esz = dbl ? 8 : 4;
	const size_t nout = (size_t) ngroup * (size_t) m->nrow;
	int32_t *g0 = normalise_groups(group, m->nleaf, ngroup);
	if (g0 == NULL)
		return SVTGPU_ERR_ARG;
	if (nout == 0) {
		free(g0);
		return SVTGPU_OK;
	}
	{   /* (not SVT_CHECK: g0 must not leak when the upload failed) */
		const int rcu = svtgpu_matrix_finish_upload(m);
		if (rcu != SVTGPU_OK) {
			free(g0);
			return rcu;
		}
	}
	if (m->nnz == 0) {
		free(g0);
		memset(out, 0, esz * nout);
		return SVTGPU_OK;
	}
	cudaStream_t s = 0;
	/* integer input: is every possible partial sum inside the int range? */
	bool small = false;
	if (!dbl && strcmp(svtgpu_env("SVTGPU_COLSUM_IMPL", "auto"),
			   "exact64") != 0) {
		int rcb = svtgpu_ensure_absmax(m, s);
		if (rcb != SVTGPU_OK) {
			free(g0);
			return rcb;
		}
		const int64_t B = svtgpu_value_bound(m);
		small = B >= 0 && (B == 0 || m->nleaf <= (int64_t) INT_MAX / B);
	}
	const size_t g_bytes = (sizeof(int32_t) * (size_t) m->nleaf + 255) &
			       ~(size_t) 255;
	const size_t cell8 = (8 * nout + 255) & ~(size_t) 255;
	const size_t cell4 = (4 * nout + 255) & ~(size_t) 255;
	/* group | overflow | out | int: sum, abs, #NA  /  double: last NA, NaN */
	/* bounded counts whose rows fit one SM's shared memory: visit the
	   leaves group by group (perm + pieces of <= piece_len leaves) */
	const bool fits32 = sizeof(int) * (size_t) m->nrow <= (size_t) 200 * 1024;
	/* more rows: 16-bit halves when the values are non-negative and a
	   piece (<= piece_len leaves, one nonzero per row and leaf) cannot
	   reach 2^16 */
	const int64_t Bv = small ? svtgpu_value_bound(m) : -1;
	int64_t piece_len = (m->nleaf + (int64_t) svtgpu_sm_count() * 4 - 1) /
			    ((int64_t) svtgpu_sm_count() * 4);
	if (piece_len < 64) piece_len = 64;
	const bool half16 = small && !fits32 && m->vmin >= 0 &&
		sizeof(short) * (size_t) (m->nrow + 1) <= (size_t) 200 * 1024 &&
		Bv >= 0 && piece_len * (Bv > 0 ? Bv : 1) <= 65535;
	const bool by_pieces = small && m->nleaf < INT_MAX &&
		(fits32 || half16) &&
		strcmp(svtgpu_env("SVTGPU_COLSUM_IMPL", "auto"), "l2") != 0;
	int32_t *h_perm = NULL;
	ColsumPiece *h_pieces = NULL;
	int npieces = 0;
	if (by_pieces) {
		const int64_t n = m->nleaf;
		h_perm = (int32_t *) malloc(sizeof(int32_t) * (size_t) n);
		int64_t *count = (int64_t *) calloc((size_t) ngroup + 1, 8);
		h_pieces = (ColsumPiece *) malloc(sizeof(ColsumPiece) *
			(size_t) (n / piece_len + ngroup + 2));
		if (h_perm == NULL || count == NULL || h_pieces == NULL) {
			free(h_perm); free(count); free(h_pieces); free(g0);
			svtgpu_set_error("out of host memory");
			return SVTGPU_ERR_NOMEM;
		}
		for (int64_t j = 0; j < n; j++)
			count[g0[j] + 1]++;
		for (int g = 0; g < ngroup; g++)
			count[g + 1] += count[g];
		for (int g = 0; g < ngroup; g++)
			for (int64_t b = count[g]; b < count[g + 1];
			     b += piece_len) {
				ColsumPiece pc;
				pc.g = g;
				pc.begin = (int32_t) b;
				pc.end = (int32_t) (b + piece_len < count[g + 1]
						    ? b + piece_len : count[g + 1]);
				pc.pad = 0;
				h_pieces[npieces++] = pc;
			}
		for (int64_t j = 0; j < n; j++)    /* stable: columns ascend */
			h_perm[count[g0[j]]++] = (int32_t) j;
		free(count);
	}
	const size_t perm_bytes = by_pieces
		? ((sizeof(int32_t) * (size_t) m->nleaf + 255) & ~(size_t) 255) : 0;
	const size_t piece_bytes = by_pieces
		? ((sizeof(ColsumPiece) * (size_t) npieces + 255) & ~(size_t) 255) : 0;
	const size_t total = g_bytes + 256 + cell8 +
			     (dbl ? 2 * cell4 : 2 * cell8 + cell4) +
			     perm_bytes + piece_bytes;
	void *scratch = NULL;
	int rc = svtgpu_scratch(m, total, &scratch);
	if (rc != SVTGPU_OK) {
		free(g0); free(h_perm); free(h_pieces);
		return rc;
	}
	char *p = (char *) scratch;
	int32_t *d_group = (int32_t *) p;            p += g_bytes;
	int32_t *d_ov = (int32_t *) p;               p += 256;
	void *d_out = p;                             p += cell8;
	GroupSumParams P;
	memset(&P, 0, sizeof(P));
	if (dbl) {
		P.acc_a = (int32_t *) p;             p += cell4;
		P.acc_b = (int32_t *) p;             p += cell4;
	} else {
		P.acc_sum = (long long *) p;         p += cell8;
		P.acc_abs = (unsigned long long *) p; p += cell8;
		P.acc_a = (int32_t *) p;             p += cell4;
	}
	int32_t *d_perm = NULL;
	ColsumPiece *d_pieces = NULL;
	if (by_pieces) {
		d_perm = (int32_t *) p;              p += perm_bytes;
		d_pieces = (ColsumPiece *) p;        p += piece_bytes;
	}
	cudaError_t e = cudaMemcpyAsync(d_group, g0, sizeof(int32_t) *
					(size_t) m->nleaf, cudaMemcpyHostToDevice, s);
	if (e == cudaSuccess && by_pieces)
		e = cudaMemcpyAsync(d_perm, h_perm, sizeof(int32_t) *
				    (size_t) m->nleaf, cudaMemcpyHostToDevice, s);
	if (e == cudaSuccess && by_pieces)
		e = cudaMemcpyAsync(d_pieces, h_pieces, sizeof(ColsumPiece) *
				    (size_t) npieces, cudaMemcpyHostToDevice, s);
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(s);
	free(g0);
	free(h_perm);
	free(h_pieces);
	SVT_CUDA(e);
	SVT_CUDA(cudaMemsetAsync(d_ov, 0, 256 + cell8, s));   /* flag + out */
	if (dbl) {
		SVT_CUDA(cudaMemsetAsync(P.acc_a, 0xFF, 2 * cell4, s));  /* -1 */
	} else {
		SVT_CUDA(cudaMemsetAsync(P.acc_sum, 0, 2 * cell8 + cell4, s));
	}
	P.offs = m->d_offs;
	P.vals = (m->flags & SVTGPU_HAS_VALS) ? m->d_vals : NULL;
	P.leaf_ptr = m->d_leaf_ptr;
	P.group = d_group;
	P.nrow = m->nrow;
	P.nleaf = m->nleaf;
	P.ngroup = ngroup;
	P.narm = narm != 0;
	P.out = d_out;
	P.overflow = d_ov;
	P.vec_ok = (((uintptr_t) m->d_offs) & 15) == 0 &&
		   strcmp(svtgpu_env("SVTGPU_COLSUM_VEC", "1"), "1") == 0;
	const bool lac = P.vals == NULL;
	const int64_t cap = (int64_t) svtgpu_sm_count() * 8;
	int64_t blocks = (m->nleaf + 7) / 8;
	if (blocks > cap) blocks = cap;
	int64_t fblocks = ((int64_t) nout + 255) / 256;
	if (fblocks > cap) fblocks = cap;
	SvtTimer t;
	SVT_CHECK(svt_timer_begin(&t, s));
	if (dbl) {
		if (lac) colsum_scatter<double, true><<<(unsigned) blocks, 256, 0, s>>>(P);
		else     colsum_scatter<double, false><<<(unsigned) blocks, 256, 0, s>>>(P);
		colsum_finish_double<<<(unsigned) fblocks, 256, 0, s>>>(P);
	} else if (small && d_perm != NULL) {
		const size_t smem = half16
			? sizeof(int) * (size_t) ((m->nrow + 1) / 2)
			: sizeof(int) * (size_t) m->nrow;
		const int grid = npieces < svtgpu_sm_count() ? npieces
							     : svtgpu_sm_count();
#define PIECES_LAUNCH(L, H) do { \
			cudaFuncSetAttribute(colsum_pieces_small<L, H>, \
				cudaFuncAttributeMaxDynamicSharedMemorySize, \
				(int) smem); \
			colsum_pieces_small<L, H><<<grid, 1024, smem, s>>>( \
				P, d_perm, d_pieces, npieces); \
		} while (0)
		if (lac && half16)       PIECES_LAUNCH(true, true);
		else if (lac)            PIECES_LAUNCH(true, false);
		else if (half16)         PIECES_LAUNCH(false, true);
		else                     PIECES_LAUNCH(false, false);
#undef PIECES_LAUNCH
		colsum_finish_small<<<(unsigned) fblocks, 256, 0, s>>>(P);
	} else if (small) {
		if (lac) colsum_scatter_small<true><<<(unsigned) blocks, 256, 0, s>>>(P);
		else     colsum_scatter_small<false><<<(unsigned) blocks, 256, 0, s>>>(P);
		colsum_finish_small<<<(unsigned) fblocks, 256, 0, s>>>(P);
	} else {
		if (lac) {
			colsum_scatter<int32_t, true><<<(unsigned) blocks, 256, 0, s>>>(P);
			colsum_finish_int<true><<<(unsigned) fblocks, 256, 0, s>>>(P);
		} else {
			colsum_scatter<int32_t, false><<<(unsigned) blocks, 256, 0, s>>>(P);
			colsum_finish_int<false><<<(unsigned) fblocks, 256, 0, s>>>(P);
		}
	}
	cudaError_t le = cudaGetLastError();
	int rc2 = svt_timer_end(&t, &m->tm.kernel_ms);
	SVT_CUDA(le);
	SVT_CHECK(rc2);
	svtgpu_count_launch(2);
	m->tm.launches = 2;
	SVT_CHECK(svt_timer_begin(&t, s));
	int32_t h_ov = 0;
	SVT_CUDA(cudaMemcpyAsync(out, d_out, esz * nout, cudaMemcpyDeviceToHost, s));
	SVT_CUDA(cudaMemcpyAsync(&h_ov, d_ov, sizeof(int32_t),
				 cudaMemcpyDeviceToHost, s));
	SVT_CHECK(svt_timer_end(&t, &m->tm.d2h_ms));
	m->tm.d2h_bytes = (double) (esz * nout);
	if (overflow != NULL)
		*overflow = h_ov != 0;
	return SVTGPU_OK;
}
