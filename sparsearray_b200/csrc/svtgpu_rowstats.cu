/* Row statistics of a device CSC -- the replacement for REC_rowStats_SVT() and
 * the update_out_for_row*() scatter loops, which the reference runs on a
 * single thread (src/SparseArray_matrixStats.c:498-829).
 *
 * Every stored value (offset, value) updates one of nrow per-row slots.  The
 * result is staged as a per-row "state" (see svt_semantics.h) which a tiny
 * kernel turns into R's answer; with a column-sharded matrix the states of the
 * shards are summed (NCCL allreduce) between the two steps.
 *
 *   row_tiles (default)  atomic-free two-pass scheme.  The grid is
 *       nchunks x ntiles CTAs.  A CTA owns the leaves of one chunk (chunks are
 *       balanced by nnz) and the rows of one tile, whose accumulators live in
 *       shared memory.  Offsets ascend strictly inside a leaf
 *       (src/leaf_utils.h:14-15), so (a) the part of a leaf that falls in a
 *       tile is one contiguous run, found once per matrix by row_split, and
 *       (b) while the consumer warps apply one run, no two threads touch the
 *       same row: plain shared-memory read-modify-write, one named barrier
 *       between runs.  A producer warp streams the runs through a ring of
 *       stages with 1-D bulk async copies (TMA) completing on mbarriers.
 *       Pass 2 (row_combine) sums the per-chunk partial vectors in a fixed
 *       order, so results do not depend on scheduling.
 *       NA/NaN are rare: they bypass the accumulators and bump per-row
 *       counters in the state with global atomics.
 *   row_flat             one thread per stored value, global atomics; used for
 *       countNAs/anyNA (which then only reads offsets of NA entries), for
 *       matrices with more rows than the tiled scheme covers, and as a
 *       cross-check (SVTGPU_ROW_IMPL=flat).
 */
#include "svtgpu_internal.h"
#include "svt_ptx.cuh"

#include <string.h>

namespace {

enum RowClass { RC_COUNT = 0, RC_SUM, RC_X2, RC_MINMAX };

inline int row_class_of(int opcode)
{
	switch (opcode) {
	    case SVTGPU_OP_ANYNA: case SVTGPU_OP_COUNTNAS: return RC_COUNT;
	    case SVTGPU_OP_SUM:                            return RC_SUM;
	    case SVTGPU_OP_CENTERED_X2_SUM:                return RC_X2;
	    case SVTGPU_OP_MIN: case SVTGPU_OP_MAX:        return RC_MINMAX;
	}
	return -1;
}

/* classification shared by every row kernel: returns 0 regular, 1 NA, 2 NaN
   and the value as a double */
__device__ __forceinline__ int classify(int32_t x, double &v)
{
	v = (double) x;
	return x == SVT_NA_INT ? 1 : 0;
}

__device__ __forceinline__ int classify(double x, double &v)
{
	v = x;
	if (!svt_isnan(x))
		return 0;
	return (uint32_t) svt_d2u(x) == 1954u ? 1 : 2;
}

__device__ __forceinline__ void atomic_min_double(double *addr, double v)
{
	unsigned long long *a = (unsigned long long *) addr;
	unsigned long long old = *a;
	while (v < __longlong_as_double((long long) old)) {
		unsigned long long assumed = old;
		old = atomicCAS(a, assumed,
				(unsigned long long) __double_as_longlong(v));
		if (old == assumed)
			break;
	}
}

__device__ __forceinline__ void atomic_max_double(double *addr, double v)
{
	unsigned long long *a = (unsigned long long *) addr;
	unsigned long long old = *a;
	while (v > __longlong_as_double((long long) old)) {
		unsigned long long assumed = old;
		old = atomicCAS(a, assumed,
				(unsigned long long) __double_as_longlong(v));
		if (old == assumed)
			break;
	}
}

/* ------------------------------------------------------------------------
 * row_flat
 */
template <int RC, typename T, bool LACUNAR>
__global__ void __launch_bounds__(256)
row_flat(const int32_t *__restrict__ offs, const T *__restrict__ vals,
	 int64_t nnz, int64_t nrow, int is_min, double *state)
{
	const int64_t stride = (int64_t) gridDim.x * blockDim.x;
	for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	     e < nnz; e += stride) {
		double v = 1.0;
		int cls = 0;
		if (!LACUNAR)
			cls = classify(vals[e], v);
		if (RC == RC_COUNT) {
			if (cls != 0)
				atomicAdd(&state[(cls == 1 ? SVT_ROW_SLOT_NA
					: SVT_ROW_SLOT_NAN) * nrow + offs[e]],
					1.0);
			continue;
		}
		const int64_t r = offs[e];
		if (RC == RC_MINMAX)
			atomicAdd(&state[SVT_ROW_SLOT_CVG * nrow + r], 1.0);
		if (cls != 0) {
			atomicAdd(&state[(cls == 1 ? SVT_ROW_SLOT_NA
					: SVT_ROW_SLOT_NAN) * nrow + r], 1.0);
			continue;
		}
		if (RC == RC_SUM || RC == RC_X2)
			atomicAdd(&state[SVT_ROW_SLOT_SUM * nrow + r], v);
		if (RC == RC_X2)
			atomicAdd(&state[SVT_ROW_SLOT_SUM2 * nrow + r], v * v);
		if (RC == RC_MINMAX) {
			if (is_min)
				atomic_min_double(&state[SVT_ROW_SLOT_EXT *
							 nrow + r], v);
			else
				atomic_max_double(&state[SVT_ROW_SLOT_EXT *
							 nrow + r], v);
		}
	}
}

__global__ void fill_doubles(double *p, int64_t n, double v)
{
	const int64_t stride = (int64_t) gridDim.x * blockDim.x;
	for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	     i < n; i += stride)
		p[i] = v;
}

/* ------------------------------------------------------------------------
 * row_split: for every leaf and every interior tile boundary b*tile_rows, the
 * leaf-relative position of the first offset >= the boundary.
 */
__global__ void __launch_bounds__(256)
row_split(const int32_t *__restrict__ offs,
	  const int64_t *__restrict__ leaf_ptr, int64_t nleaf, int ntiles,
	  int tile_rows, int32_t *__restrict__ split)
{
	const int64_t total = nleaf * (ntiles - 1);
	const int64_t stride = (int64_t) gridDim.x * blockDim.x;
	for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	     i < total; i += stride) {
		const int64_t b = i / nleaf, leaf = i - b * nleaf;
		const int64_t start = leaf_ptr[leaf];
		const int32_t bound = (int32_t) ((b + 1) * tile_rows);
		int32_t lo = 0, hi = (int32_t) (leaf_ptr[leaf + 1] - start);
		while (lo < hi) {
			int32_t mid = lo + ((hi - lo) >> 1);
			if (offs[start + mid] < bound) lo = mid + 1;
			else                           hi = mid;
		}
		split[i] = lo;
	}
}

/* ------------------------------------------------------------------------
 * row_tiles
 */

#define ROW_NS 6            /* ring depth */
#define ROW_CONSUMERS 256   /* consumer threads (8 warps) */
#define ROW_THREADS (ROW_CONSUMERS + 32)

struct RowTileParams {
	const int32_t *offs;
	const void *vals;          /* NULL: lacunar */
	const int64_t *leaf_ptr;
	const int32_t *split;      /* NULL when ntiles == 1 */
	int64_t nleaf, nnz, nrow;
	int ntiles, tile_rows, nchunks;
	int stage_elems;           /* SE */
	int is_min;
	double *part;              /* [nchunks][nacc][nrow] */
	double *state;             /* NA / NaN counters (global atomics) */
};

struct __align__(16) RowItem {
	int64_t lo, hi;      /* element range of the run */
	int32_t obase;       /* element index at byte 0 of the offs stage */
	int32_t vbase_delta; /* lo - (element index at byte 0 of vals stage) */
	int32_t odelta;      /* lo - obase */
	int32_t stop;
};

__device__ __forceinline__ void consumer_barrier(void)
{
	asm volatile("bar.sync 1, %0;" :: "n"(ROW_CONSUMERS) : "memory");
}

/* smem: acc[NACC][tile_rows] doubles | ROW_NS x (offs stage | vals stage) |
 * RowItem[ROW_NS] | full[ROW_NS], empty[ROW_NS] */
template <int RC, typename T, bool LACUNAR>
__global__ void __launch_bounds__(ROW_THREADS, 1)
row_tiles(RowTileParams P)
{
	constexpr int NACC = RC == RC_SUM ? 1 : 2;
	constexpr int VSZ = (int) sizeof(T);   /* T is int32_t when LACUNAR */
	extern __shared__ __align__(128) unsigned char smem[];
	double *acc = (double *) smem;
	const int offs_stage_bytes = P.stage_elems * 4 + 32;
	const int vals_stage_bytes = LACUNAR ? 0 : P.stage_elems * VSZ + 32;
	const int stage_bytes = offs_stage_bytes + vals_stage_bytes;
	unsigned char *ring = smem + (size_t) NACC * P.tile_rows * 8;
	RowItem *items = (RowItem *) (ring + (size_t) ROW_NS * stage_bytes);
	uint64_t *bars = (uint64_t *) (items + ROW_NS);
	const uint32_t full0 = svt_smem_u32(&bars[0]);
	const uint32_t empty0 = svt_smem_u32(&bars[ROW_NS]);

	const int chunk = blockIdx.x / P.ntiles;
	const int tile = blockIdx.x - chunk * P.ntiles;
	const int row0 = tile * P.tile_rows;
	int rows_here = (int) (P.nrow - row0 < P.tile_rows ? P.nrow - row0
							    : P.tile_rows);
	if (rows_here < 0) rows_here = 0;

	if (threadIdx.x == 0) {
		for (int i = 0; i < 2 * ROW_NS; i++)
			svt_mbar_init(svt_smem_u32(&bars[i]), 1);
		svt_mbar_init_fence();
	}
	__syncthreads();

	if (threadIdx.x >= ROW_CONSUMERS) {
		/* ---- producer warp ---- */
		const int lane = threadIdx.x & 31;
		/* leaves [l0, l1) of this chunk: boundaries are the first
		   leaves starting at or after c * nnz / nchunks */
		int64_t bounds[2];
		for (int k = 0; k < 2; k++) {
			const int c = chunk + k;
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
		const int64_t l0 = bounds[0], l1 = bounds[1];
		const uint64_t policy = svt_policy_evict_first();
		uint32_t it = 0;
		for (int64_t base = l0; base < l1; base += 32) {
			/* each lane fetches the run of one leaf */
			const int64_t leaf = base + lane;
			int64_t my_lo = 0, my_hi = 0;
			if (leaf < l1) {
				const int64_t start = P.leaf_ptr[leaf];
				const int64_t nz = P.leaf_ptr[leaf + 1] - start;
				my_lo = start + (tile == 0 ? 0
					: P.split[(int64_t) (tile - 1) *
						  P.nleaf + leaf]);
				my_hi = start + (tile == P.ntiles - 1 ? nz
					: P.split[(int64_t) tile * P.nleaf +
						  leaf]);
			}
			const int n = (int) (l1 - base < 32 ? l1 - base : 32);
			for (int i = 0; i < n; i++) {
				const int64_t lo = __shfl_sync(SVT_FULL_MASK,
							       my_lo, i);
				const int64_t hi = __shfl_sync(SVT_FULL_MASK,
							       my_hi, i);
				if (lane != 0)
					continue;
				for (int64_t c0 = lo; c0 < hi;
				     c0 += P.stage_elems) {
					int64_t c1 = c0 + P.stage_elems;
					if (c1 > hi) c1 = hi;
					const int st = it % ROW_NS;
					svt_mbar_wait(empty0 + 8 * st,
						((it / ROW_NS) & 1) ^ 1);
					unsigned char *sb = ring +
						(size_t) st * stage_bytes;
					const int64_t oa0 = (c0 * 4) &
							    ~(int64_t) 15;
					const int64_t oa1 = (c1 * 4 + 15) &
							    ~(int64_t) 15;
					uint32_t bytes = (uint32_t) (oa1 - oa0);
					int64_t va0 = 0, va1 = 0;
					if (!LACUNAR) {
						va0 = (c0 * VSZ) & ~(int64_t) 15;
						va1 = (c1 * VSZ + 15) &
						      ~(int64_t) 15;
						bytes += (uint32_t) (va1 - va0);
					}
					RowItem ri;
					ri.lo = c0;
					ri.hi = c1;
					ri.obase = 0;
					ri.odelta = (int32_t) (c0 - oa0 / 4);
					ri.vbase_delta = LACUNAR ? 0
						: (int32_t) (c0 - va0 / VSZ);
					ri.stop = 0;
					items[st] = ri;
					svt_mbar_arrive_expect_tx(
						full0 + 8 * st, bytes);
					svt_bulk_g2s_hint(svt_smem_u32(sb),
						(const char *) P.offs + oa0,
						(uint32_t) (oa1 - oa0),
						full0 + 8 * st, policy);
					if (!LACUNAR)
						svt_bulk_g2s_hint(
						    svt_smem_u32(sb +
							offs_stage_bytes),
						    (const char *) P.vals + va0,
						    (uint32_t) (va1 - va0),
						    full0 + 8 * st, policy);
					it++;
				}
			}
			__syncwarp();
		}
		if (lane == 0) {
			const int st = it % ROW_NS;
			svt_mbar_wait(empty0 + 8 * st, ((it / ROW_NS) & 1) ^ 1);
			RowItem ri;
			ri.lo = ri.hi = 0; ri.obase = ri.odelta = 0;
			ri.vbase_delta = 0; ri.stop = 1;
			items[st] = ri;
			svt_mbar_arrive(full0 + 8 * st);
		}
		return;
	}

	/* ---- consumer warps ---- */
	const int ctid = threadIdx.x;
	double *acc0 = acc;                    /* sum | coverage */
	double *acc1 = acc + P.tile_rows;      /* sum2 | extreme */
	const double ext_init = P.is_min ? svt_posinf() : svt_neginf();
	for (int r = ctid; r < P.tile_rows; r += ROW_CONSUMERS) {
		acc0[r] = 0.0;
		if (NACC == 2)
			acc1[r] = RC == RC_MINMAX ? ext_init : 0.0;
	}
	consumer_barrier();

	for (uint32_t it = 0; ; it++) {
		const int st = it % ROW_NS;
		svt_mbar_wait(full0 + 8 * st, (it / ROW_NS) & 1);
		const RowItem ri = items[st];
		if (ri.stop)
			break;
		const unsigned char *sb = ring + (size_t) st * stage_bytes;
		const int32_t *so = (const int32_t *) sb + ri.odelta;
		const T *sv = (const T *) (sb + offs_stage_bytes) +
			      ri.vbase_delta;
		const int n = (int) (ri.hi - ri.lo);
		for (int e = ctid; e < n; e += ROW_CONSUMERS) {
			const int off = so[e];
			const int r = off - row0;
			double v = 1.0;
			int cls = 0;
			if (!LACUNAR)
				cls = classify(sv[e], v);
			if (RC == RC_MINMAX)
				acc0[r] += 1.0;
			if (cls != 0) {
				atomicAdd(&P.state[(cls == 1 ? SVT_ROW_SLOT_NA
					: SVT_ROW_SLOT_NAN) * P.nrow + off],
					1.0);
				continue;
			}
			if (RC == RC_SUM || RC == RC_X2)
				acc0[r] += v;
			if (RC == RC_X2)
				acc1[r] += v * v;
			if (RC == RC_MINMAX) {
				const double cur = acc1[r];
				if (P.is_min ? v < cur : v > cur)
					acc1[r] = v;
			}
		}
		/* all updates of this run are done before the next one (which
		   may hit the same rows) starts, and the stage can be reused */
		consumer_barrier();
		if (ctid == 0)
			svt_mbar_arrive(empty0 + 8 * st);
	}

	/* flush this CTA's partial rows */
	double *part = P.part + (size_t) chunk * NACC * P.nrow;
	for (int r = ctid; r < rows_here; r += ROW_CONSUMERS) {
		part[row0 + r] = acc0[r];
		if (NACC == 2)
			part[P.nrow + row0 + r] = acc1[r];
	}
}

/* pass 2: fixed-order sum (or min/max) of the per-chunk partial vectors into
 * the state slots */
template <int RC>
__global__ void __launch_bounds__(256)
row_combine(const double *__restrict__ part, int nchunks, int64_t nrow,
	    int is_min, double *__restrict__ state)
{
	constexpr int NACC = RC == RC_SUM ? 1 : 2;
	const int64_t stride = (int64_t) gridDim.x * blockDim.x;
	for (int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	     r < nrow; r += stride) {
		double a0 = 0.0;
		double a1 = RC == RC_MINMAX
			? (is_min ? svt_posinf() : svt_neginf()) : 0.0;
		for (int c = 0; c < nchunks; c++) {
			const double *p = part + (size_t) c * NACC * nrow;
			a0 += p[r];
			if (NACC == 2) {
				const double x = p[nrow + r];
				if (RC == RC_MINMAX)
					a1 = (is_min ? x < a1 : x > a1) ? x : a1;
				else
					a1 += x;
			}
		}
		if (RC == RC_SUM) {
			state[SVT_ROW_SLOT_SUM * nrow + r] = a0;
		} else if (RC == RC_X2) {
			state[SVT_ROW_SLOT_SUM * nrow + r] = a0;
			state[SVT_ROW_SLOT_SUM2 * nrow + r] = a1;
		} else {
			state[SVT_ROW_SLOT_CVG * nrow + r] = a0;
			state[SVT_ROW_SLOT_EXT * nrow + r] = a1;
		}
	}
}

/* state -> R's answer */
__global__ void __launch_bounds__(256)
row_finalize(int opcode, int is_double, int narm, int64_t nrow,
	     int64_t nstrata, const double *__restrict__ center,
	     const double *__restrict__ state, void *out, int32_t *warn)
{
	const int out_is_int = opcode == SVTGPU_OP_ANYNA ||
		((opcode == SVTGPU_OP_MIN || opcode == SVTGPU_OP_MAX) &&
		 !is_double);
	const int64_t stride = (int64_t) gridDim.x * blockDim.x;
	for (int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	     r < nrow; r += stride) {
		SvtScalar s = svt_row_finalize(opcode, is_double, narm,
				nstrata, center != NULL,
				center != NULL ? center[r] : 0.0,
				state + r, nrow);
		if (out_is_int) ((int32_t *) out)[r] = s.i;
		else            ((double *) out)[r] = s.d;
		if (s.warn && warn != NULL)
			atomicOr((int *) warn, 1);
	}
}

__global__ void __launch_bounds__(256)
row_moments_finalize(int narm, int64_t nrow, int64_t nstrata,
		     const double *__restrict__ state, double *mean,
		     double *var)
{
	const int64_t stride = (int64_t) gridDim.x * blockDim.x;
	for (int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	     r < nrow; r += stride) {
		double mu, v;
		svt_row_moments(narm, nstrata, state + r, nrow, &mu, &v);
		if (mean != NULL) mean[r] = mu;
		if (var != NULL)  var[r] = v;
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

struct TileConfig {
	int ok;
	int ntiles, tile_rows, nchunks, stage_elems;
	size_t smem;
};

/* smallest number of row tiles whose accumulators + staging ring fit */
TileConfig choose_tiles(int64_t nrow, int64_t nleaf, int64_t nnz, int nacc,
			int vsz)
{
	TileConfig c;
	memset(&c, 0, sizeof(c));
	const size_t budget = 222 * 1024;
	const int sms = svtgpu_sm_count();
	const double avg_leaf = nleaf > 0 ? (double) nnz / (double) nleaf : 0.0;
	const int force = atoi(svtgpu_env("SVTGPU_ROW_NTILES", "0"));
	for (int nt = force > 0 ? force : 1; nt <= 64; nt++) {
		int64_t tr = (nrow + nt - 1) / nt;
		tr = (tr + 31) / 32 * 32;
		if (tr < 32) tr = 32;
		double item = avg_leaf / nt;
		int se = (int) (item * 1.2) + 64;
		se = (se + 127) / 128 * 128;
		if (se < 256) se = 256;
		if (se > 4096) se = 4096;
		size_t smem = (size_t) nacc * tr * 8 +
			(size_t) ROW_NS * ((size_t) se * 4 + 32 +
				(vsz ? (size_t) se * vsz + 32 : 0)) +
			ROW_NS * sizeof(RowItem) + 2 * ROW_NS * 8 + 128;
		if (smem <= budget) {
			c.ok = 1;
			c.ntiles = nt;
			c.tile_rows = (int) tr;
			c.stage_elems = se;
			c.smem = smem;
			c.nchunks = sms / nt;
			if (c.nchunks < 1) c.nchunks = 1;
			if ((int64_t) c.nchunks > nleaf)
				c.nchunks = nleaf > 0 ? (int) nleaf : 1;
			return c;
		}
		if (force > 0)
			break;
	}
	return c;
}

int ensure_split(svtgpu_matrix *m, const TileConfig &c, cudaStream_t s)
{
	if (c.ntiles <= 1)
		return SVTGPU_OK;
	if (m->d_split != NULL && m->split_tile_rows == c.tile_rows &&
	    m->split_ntiles == c.ntiles)
		return SVTGPU_OK;
	if (m->d_split != NULL) {
		SVT_CUDA(cudaStreamSynchronize(s));
		SVT_CUDA(cudaFree(m->d_split));
		m->d_split = NULL;
	}
	const int64_t n = m->nleaf * (c.ntiles - 1);
	SVT_CUDA(cudaMalloc((void **) &m->d_split,
			    sizeof(int32_t) * (size_t) (n > 0 ? n : 1)));
	row_split<<<grid_for(n, 256), 256, 0, s>>>(m->d_offs, m->d_leaf_ptr,
			m->nleaf, c.ntiles, c.tile_rows, m->d_split);
	SVT_CUDA(cudaGetLastError());
	svtgpu_count_launch(1);
	m->split_tile_rows = c.tile_rows;
	m->split_ntiles = c.ntiles;
	return SVTGPU_OK;
}

template <int RC, typename T, bool LAC>
int launch_flat(const svtgpu_matrix *m, int is_min, double *d_state,
		cudaStream_t s)
{
	row_flat<RC, T, LAC><<<grid_for(m->nnz, 256 * 8), 256, 0, s>>>(
		m->d_offs, (const T *) m->d_vals, m->nnz, m->nrow, is_min,
		d_state);
	SVT_CUDA(cudaGetLastError());
	svtgpu_count_launch(1);
	return SVTGPU_OK;
}

template <int RC, typename T, bool LAC>
int launch_tiles(svtgpu_matrix *m, const TileConfig &c, int is_min,
		 double *d_state, cudaStream_t s)
{
	constexpr int NACC = RC == RC_SUM ? 1 : 2;
	SVT_CHECK(ensure_split(m, c, s));
	void *part = NULL;
	SVT_CHECK(svtgpu_scratch(m, sizeof(double) * (size_t) c.nchunks *
				 NACC * (size_t) m->nrow + 64, &part));
	RowTileParams P;
	P.offs = m->d_offs;
	P.vals = LAC ? NULL : m->d_vals;
	P.leaf_ptr = m->d_leaf_ptr;
	P.split = m->d_split;
	P.nleaf = m->nleaf;
	P.nnz = m->nnz;
	P.nrow = m->nrow;
	P.ntiles = c.ntiles;
	P.tile_rows = c.tile_rows;
	P.nchunks = c.nchunks;
	P.stage_elems = c.stage_elems;
	P.is_min = is_min;
	P.part = (double *) part;
	P.state = d_state;
	SVT_CUDA(cudaFuncSetAttribute(row_tiles<RC, T, LAC>,
		cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c.smem));
	row_tiles<RC, T, LAC><<<(unsigned) (c.nchunks * c.ntiles), ROW_THREADS,
				c.smem, s>>>(P);
	SVT_CUDA(cudaGetLastError());
	row_combine<RC><<<grid_for(m->nrow, 256), 256, 0, s>>>(
		(const double *) part, c.nchunks, m->nrow, is_min, d_state);
	SVT_CUDA(cudaGetLastError());
	svtgpu_count_launch(2);
	return SVTGPU_OK;
}

template <int RC>
int launch_class(svtgpu_matrix *m, bool tiles, const TileConfig &c,
		 int is_min, double *d_state, cudaStream_t s)
{
	const bool lac = !(m->flags & SVTGPU_HAS_VALS);
	const bool dbl = svt_is_double(m->val_type);
	if (tiles) {
		if (lac) return launch_tiles<RC, int32_t, true>(m, c, is_min,
								d_state, s);
		if (dbl) return launch_tiles<RC, double, false>(m, c, is_min,
								d_state, s);
		return launch_tiles<RC, int32_t, false>(m, c, is_min, d_state, s);
	}
	if (lac) return launch_flat<RC, int32_t, true>(m, is_min, d_state, s);
	if (dbl) return launch_flat<RC, double, false>(m, is_min, d_state, s);
	return launch_flat<RC, int32_t, false>(m, is_min, d_state, s);
}

}  /* namespace */

/* Reduce the leaves of `m` into a state of (n_sum + n_ext) x nrow doubles.
 * want_sum2 upgrades SUM to the {sum, sum2} accumulation used by rowVars. */
int svtgpu_launch_row_accumulate(svtgpu_matrix *m, int opcode, int narm,
				 int want_sum2, double *d_state,
				 cudaStream_t s)
{
	(void) narm;   /* NA handling is decided when the state is finalised */
	SVT_ARG(svt_row_op_supported(opcode),
		"rowStats: operation %d is not supported natively (the "
		"reference only implements countNAs, anyNA, min, max, sum and "
		"centered_X2_sum in C_rowStats_SVT)", opcode);
	SVT_ARG((m->flags & SVTGPU_HAS_OFFS) || m->nnz == 0,
		"rowStats: the matrix was uploaded without row offsets");
	int rc_class = row_class_of(opcode);
	if (rc_class == RC_SUM && want_sum2)
		rc_class = RC_X2;
	const int is_min = opcode == SVTGPU_OP_MIN;
	int n_sum = 0, n_ext = 0;
	svt_row_state_layout(rc_class == RC_X2 ? SVTGPU_OP_CENTERED_X2_SUM
					       : opcode, &n_sum, &n_ext);
	const int64_t nrow = m->nrow;
	if (nrow == 0)
		return SVTGPU_OK;
	SVT_CUDA(cudaMemsetAsync(d_state, 0,
				 sizeof(double) * (size_t) (n_sum * nrow), s));
	if (n_ext > 0) {
		fill_doubles<<<grid_for(nrow, 256), 256, 0, s>>>(
			d_state + (size_t) n_sum * nrow, nrow,
			is_min ? svt_posinf() : svt_neginf());
		SVT_CUDA(cudaGetLastError());
		svtgpu_count_launch(1);
	}
	if (m->nnz == 0)
		return SVTGPU_OK;
	if (rc_class == RC_COUNT) {
		/* lacunar leaves hold no NA: nothing to scan */
		if (!(m->flags & SVTGPU_HAS_VALS))
			return SVTGPU_OK;
		TileConfig none;
		memset(&none, 0, sizeof(none));
		return launch_class<RC_COUNT>(m, false, none, 0, d_state, s);
	}
	const int nacc = rc_class == RC_SUM ? 1 : 2;
	const int vsz = (m->flags & SVTGPU_HAS_VALS)
			? (int) svt_val_size(m->val_type) : 0;
	TileConfig c = choose_tiles(nrow, m->nleaf, m->nnz, nacc, vsz);
	const char *impl = svtgpu_env("SVTGPU_ROW_IMPL", "tiles");
	const bool aligned = (((uintptr_t) m->d_offs) & 15) == 0 &&
			     (((uintptr_t) m->d_vals) & 15) == 0;
	const bool tiles = c.ok && aligned && strcmp(impl, "flat") != 0;
	switch (rc_class) {
	    case RC_SUM:
		return launch_class<RC_SUM>(m, tiles, c, is_min, d_state, s);
	    case RC_X2:
		return launch_class<RC_X2>(m, tiles, c, is_min, d_state, s);
	    case RC_MINMAX:
		return launch_class<RC_MINMAX>(m, tiles, c, is_min, d_state, s);
	}
	svtgpu_set_error("rowStats: internal error (row class)");
	return SVTGPU_ERR_ARG;
}

int svtgpu_launch_row_finalize(int opcode, int val_type, int narm,
			       int64_t nrow, int64_t nstrata,
			       const double *d_center, const double *d_state,
			       void *d_out, int32_t *d_warn, cudaStream_t s)
{
	if (nrow == 0)
		return SVTGPU_OK;
	row_finalize<<<grid_for(nrow, 256), 256, 0, s>>>(opcode,
		svt_is_double(val_type), narm != 0, nrow, nstrata, d_center,
		d_state, d_out, d_warn);
	SVT_CUDA(cudaGetLastError());
	svtgpu_count_launch(1);
	return SVTGPU_OK;
}

int svtgpu_launch_row_moments_finalize(int val_type, int narm, int64_t nrow,
				       int64_t nstrata, const double *d_state,
				       double *d_mean, double *d_var,
				       cudaStream_t s)
{
	(void) val_type;
	if (nrow == 0)
		return SVTGPU_OK;
	row_moments_finalize<<<grid_for(nrow, 256), 256, 0, s>>>(narm != 0,
		nrow, nstrata, d_state, d_mean, d_var);
	SVT_CUDA(cudaGetLastError());
	svtgpu_count_launch(1);
	return SVTGPU_OK;
}

/* ---- C ABI ---- */

extern "C" int svtgpu_rowstats_state_layout(int opcode, int val_type,
					    int *n_sum_slots,
					    int *n_minmax_slots)
{
	(void) val_type;
	SVT_ARG(svt_row_op_supported(opcode),
		"rowStats: unsupported operation %d", opcode);
	svt_row_state_layout(opcode, n_sum_slots, n_minmax_slots);
	return SVTGPU_OK;
}

extern "C" int svtgpu_rowstats_accumulate_dev(svtgpu_matrix *m, int opcode,
					      int narm, double *d_state,
					      void *stream)
{
	SVT_CHECK(svtgpu_require_device());
	SVT_ARG(m != NULL && d_state != NULL,
		"svtgpu_rowstats_accumulate_dev: NULL argument");
	return svtgpu_launch_row_accumulate(m, opcode, narm, 0, d_state,
					    (cudaStream_t) stream);
}

extern "C" int svtgpu_rowstats_finalize_dev(int opcode, int val_type,
		int narm, int64_t nrow, int64_t nstrata_total,
		const double *d_center, const double *d_state, void *d_out,
		int32_t *d_warn, void *stream)
{
	SVT_CHECK(svtgpu_require_device());
	SVT_ARG(svt_row_op_supported(opcode),
		"rowStats: unsupported operation %d", opcode);
	return svtgpu_launch_row_finalize(opcode, val_type, narm, nrow,
			nstrata_total, d_center, d_state, d_out, d_warn,
			(cudaStream_t) stream);
}

extern "C" int svtgpu_rowmoments_accumulate_dev(svtgpu_matrix *m, int narm,
						double *d_state, void *stream)
{
	SVT_CHECK(svtgpu_require_device());
	SVT_ARG(m != NULL && d_state != NULL,
		"svtgpu_rowmoments_accumulate_dev: NULL argument");
	return svtgpu_launch_row_accumulate(m, SVTGPU_OP_SUM, narm, 1, d_state,
					    (cudaStream_t) stream);
}

extern "C" int svtgpu_rowmoments_finalize_dev(int val_type, int narm,
		int64_t nrow, int64_t nstrata_total, const double *d_state,
		double *d_mean, double *d_var, void *stream)
{
	SVT_CHECK(svtgpu_require_device());
	return svtgpu_launch_row_moments_finalize(val_type, narm, nrow,
			nstrata_total, d_state, d_mean, d_var,
			(cudaStream_t) stream);
}

extern "C" int svtgpu_rowstats(svtgpu_matrix *m, int opcode, int narm,
			       const double *center, void *out, int *warn)
{
	SVT_CHECK(svtgpu_require_device());
	SVT_ARG(m != NULL && out != NULL, "svtgpu_rowstats: NULL argument");
	SVT_ARG(svt_row_op_supported(opcode),
		"rowStats: operation %d is not supported natively (the "
		"reference only implements countNAs, anyNA, min, max, sum and "
		"centered_X2_sum in C_rowStats_SVT)", opcode);
	SVT_CHECK(svtgpu_matrix_finish_upload(m));
	if (warn != NULL)
		*warn = 0;
	m->tm.kernel_ms = m->tm.d2h_ms = 0.0;
	m->tm.d2h_bytes = 0.0;
	m->tm.launches = 0;
	const int64_t nrow = m->nrow;
	if (nrow == 0)
		return SVTGPU_OK;
	const int out_is_int = opcode == SVTGPU_OP_ANYNA ||
		((opcode == SVTGPU_OP_MIN || opcode == SVTGPU_OP_MAX) &&
		 !svt_is_double(m->val_type));
	const size_t esz = out_is_int ? 4 : 8;
	cudaStream_t s = 0;
	/* state (4 slots) | out | center | warn in one side allocation: the
	   matrix scratch is used by the accumulate step for partials */
	double *d_buf = NULL;
	SVT_CUDA(cudaMalloc((void **) &d_buf,
			    sizeof(double) * (size_t) (6 * nrow) + 64));
	double *d_state = d_buf;
	void *d_out = d_buf + 4 * nrow;
	double *d_center = d_buf + 5 * nrow;
	int32_t *d_warn = (int32_t *) (d_buf + 6 * nrow);
	int rc = SVTGPU_OK;
	const bool use_center = center != NULL &&
				opcode == SVTGPU_OP_CENTERED_X2_SUM;
	cudaError_t e = cudaMemsetAsync(d_warn, 0, 64, s);
	if (!use_center)
		d_center = NULL;
	else if (e == cudaSuccess)
		e = cudaMemcpyAsync(d_center, center, sizeof(double) * nrow,
				    cudaMemcpyHostToDevice, s);
	if (e != cudaSuccess) {
		rc = svtgpu_cuda_fail(e, "svtgpu_rowstats setup", __FILE__,
				      __LINE__);
		cudaFree(d_buf);
		return rc;
	}
	SvtTimer t;
	int64_t l0 = svtgpu_launch_count();
	rc = svt_timer_begin(&t, s);
	if (rc == SVTGPU_OK)
		rc = svtgpu_launch_row_accumulate(m, opcode, narm, 0, d_state,
						  s);
	if (rc == SVTGPU_OK)
		rc = svtgpu_launch_row_finalize(opcode, m->val_type, narm,
				nrow, m->nleaf, d_center, d_state, d_out,
				d_warn, s);
	int rc2 = svt_timer_end(&t, &m->tm.kernel_ms);
	if (rc == SVTGPU_OK)
		rc = rc2;
	m->tm.launches = (int) (svtgpu_launch_count() - l0);
	int32_t h_warn = 0;
	if (rc == SVTGPU_OK) {
		rc = svt_timer_begin(&t, s);
		e = cudaMemcpyAsync(out, d_out, esz * (size_t) nrow,
				    cudaMemcpyDeviceToHost, s);
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(&h_warn, d_warn, sizeof(int32_t),
					    cudaMemcpyDeviceToHost, s);
		rc2 = svt_timer_end(&t, &m->tm.d2h_ms);
		if (e != cudaSuccess)
			rc = svtgpu_cuda_fail(e, "svtgpu_rowstats D2H",
					      __FILE__, __LINE__);
		else if (rc == SVTGPU_OK)
			rc = rc2;
		m->tm.d2h_bytes = (double) (esz * (size_t) nrow);
	}
	cudaFree(d_buf);
	if (rc == SVTGPU_OK && warn != NULL)
		*warn = h_warn != 0;
	return rc;
}

extern "C" int svtgpu_rowmoments(svtgpu_matrix *m, int narm, double *out_mean,
				 double *out_var)
{
	SVT_CHECK(svtgpu_require_device());
	SVT_ARG(m != NULL, "svtgpu_rowmoments: NULL matrix");
	SVT_CHECK(svtgpu_matrix_finish_upload(m));
	m->tm.kernel_ms = m->tm.d2h_ms = 0.0;
	m->tm.d2h_bytes = 0.0;
	m->tm.launches = 0;
	const int64_t nrow = m->nrow;
	if (nrow == 0)
		return SVTGPU_OK;
	cudaStream_t s = 0;
	double *d_buf = NULL;
	SVT_CUDA(cudaMalloc((void **) &d_buf,
			    sizeof(double) * (size_t) (6 * nrow)));
	double *d_state = d_buf, *d_mean = d_buf + 4 * nrow,
	       *d_var = d_buf + 5 * nrow;
	SvtTimer t;
	int64_t l0 = svtgpu_launch_count();
	int rc = svt_timer_begin(&t, s);
	if (rc == SVTGPU_OK)
		rc = svtgpu_launch_row_accumulate(m, SVTGPU_OP_SUM, narm, 1,
						  d_state, s);
	if (rc == SVTGPU_OK)
		rc = svtgpu_launch_row_moments_finalize(m->val_type, narm,
				nrow, m->nleaf, d_state, d_mean, d_var, s);
	int rc2 = svt_timer_end(&t, &m->tm.kernel_ms);
	if (rc == SVTGPU_OK)
		rc = rc2;
	m->tm.launches = (int) (svtgpu_launch_count() - l0);
	if (rc == SVTGPU_OK) {
		rc = svt_timer_begin(&t, s);
		cudaError_t e = cudaSuccess;
		if (out_mean != NULL)
			e = cudaMemcpyAsync(out_mean, d_mean,
					    sizeof(double) * nrow,
					    cudaMemcpyDeviceToHost, s);
		if (e == cudaSuccess && out_var != NULL)
			e = cudaMemcpyAsync(out_var, d_var,
					    sizeof(double) * nrow,
					    cudaMemcpyDeviceToHost, s);
		rc2 = svt_timer_end(&t, &m->tm.d2h_ms);
		if (e != cudaSuccess)
			rc = svtgpu_cuda_fail(e, "svtgpu_rowmoments D2H",
					      __FILE__, __LINE__);
		else if (rc == SVTGPU_OK)
			rc = rc2;
		m->tm.d2h_bytes = 8.0 * (double) nrow *
			((out_mean != NULL) + (out_var != NULL));
	}
	cudaFree(d_buf);
	return rc;
}
