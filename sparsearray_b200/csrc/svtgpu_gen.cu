/* Synthetic SVT generation directly in HBM (benchmark inputs only; see
 * include/svtgpu.h for the counter-based formula and its reference). */
#include "svtgpu_internal.h"
#include "svt_ptx.cuh"

namespace {

#define GEN_MAX_THRESH 16

struct GenParams {
	int64_t nrow, nleaf, leaf0;
	uint64_t seed;
	uint32_t nz_threshold, na_threshold;
	uint32_t vt[GEN_MAX_THRESH];
	int nvt;
};

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
	z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
	z ^= z >> 27; z *= 0x94D049BB133111EBULL;
	z ^= z >> 31;
	return z;
}

__device__ __forceinline__ uint64_t cell_hash(const GenParams &G,
					      int64_t leaf, int64_t i)
{
	const uint64_t cell = (uint64_t) (G.leaf0 + leaf) * (uint64_t) G.nrow +
			      (uint64_t) i;
	return mix64(G.seed + (cell + 1) * 0x9E3779B97F4A7C15ULL);
}

__global__ void __launch_bounds__(256)
gen_count(GenParams G, int64_t *leaf_nnz)
{
	const int lane = threadIdx.x & 31;
	const int64_t leaf = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (leaf >= G.nleaf)
		return;
	long long c = 0;
	for (int64_t i = lane; i < G.nrow; i += 32)
		c += (uint32_t) (cell_hash(G, leaf, i) >> 32) < G.nz_threshold;
	c = svt_warp_sum(c);
	if (lane == 0)
		leaf_nnz[leaf] = c;
}

template <typename T, bool HAS_VALS>
__global__ void __launch_bounds__(256)
gen_fill(GenParams G, const int64_t *__restrict__ leaf_ptr,
	 int32_t *__restrict__ offs, T *__restrict__ vals)
{
	const int lane = threadIdx.x & 31;
	const int64_t leaf = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (leaf >= G.nleaf)
		return;
	int64_t base = leaf_ptr[leaf];
	for (int64_t i0 = 0; i0 < G.nrow; i0 += 32) {
		const int64_t i = i0 + lane;
		uint64_t h = 0;
		bool hit = false;
		if (i < G.nrow) {
			h = cell_hash(G, leaf, i);
			hit = (uint32_t) (h >> 32) < G.nz_threshold;
		}
		const unsigned mask = __ballot_sync(SVT_FULL_MASK, hit);
		if (hit) {
			const int64_t pos = base +
				__popc(mask & ((1u << lane) - 1u));
			offs[pos] = (int32_t) i;
			if (HAS_VALS) {
				const uint64_t h2 = mix64(h ^
						0xD1B54A32D192ED03ULL);
				const uint32_t u = (uint32_t) (h2 >> 32);
				int v = 1;
				for (int j = 0; j < G.nvt; j++)
					v += u >= G.vt[j];
				if ((uint32_t) h2 < G.na_threshold) {
					if (sizeof(T) == 4)
						vals[pos] = (T) SVT_NA_INT;
					else
						vals[pos] = (T) svt_na_real();
				} else {
					vals[pos] = (T) v;
				}
			}
		}
		base += __popc(mask);
	}
}

int fill_params(GenParams *G, int64_t nrow, int64_t nleaf, int64_t leaf0,
		uint64_t seed, uint32_t nz_threshold, uint32_t na_threshold,
		const uint32_t *vt, int nvt)
{
	SVT_ARG(nrow >= 0 && nrow <= INT32_MAX && nleaf >= 0,
		"svtgpu_gen: bad extents");
	SVT_ARG(nvt >= 0 && nvt <= GEN_MAX_THRESH,
		"svtgpu_gen: at most %d value thresholds", GEN_MAX_THRESH);
	G->nrow = nrow; G->nleaf = nleaf; G->leaf0 = leaf0; G->seed = seed;
	G->nz_threshold = nz_threshold; G->na_threshold = na_threshold;
	G->nvt = nvt;
	for (int j = 0; j < GEN_MAX_THRESH; j++)
		G->vt[j] = j < nvt ? vt[j] : 0xFFFFFFFFu;
	return SVTGPU_OK;
}

}  /* namespace */

extern "C" int svtgpu_gen_count(int64_t nrow, int64_t nleaf, int64_t leaf0,
				uint64_t seed, uint32_t nz_threshold,
				int64_t *d_leaf_nnz, void *stream)
{
	SVT_CHECK(svtgpu_require_device());
	GenParams G;
	SVT_CHECK(fill_params(&G, nrow, nleaf, leaf0, seed, nz_threshold, 0,
			      NULL, 0));
	if (nleaf == 0)
		return SVTGPU_OK;
	const int64_t blocks = (nleaf + 7) / 8;
	SVT_ARG(blocks <= INT32_MAX, "svtgpu_gen_count: too many leaves");
	gen_count<<<(unsigned) blocks, 256, 0, (cudaStream_t) stream>>>(
		G, d_leaf_nnz);
	SVT_CUDA(cudaGetLastError());
	svtgpu_count_launch(1);
	return SVTGPU_OK;
}

extern "C" int svtgpu_gen_fill(int64_t nrow, int64_t nleaf, int64_t leaf0,
			       uint64_t seed, uint32_t nz_threshold,
			       uint32_t na_threshold,
			       const uint32_t *value_thresholds,
			       int n_value_thresholds, int val_type,
			       const int64_t *d_leaf_ptr, int32_t *d_offs,
			       void *d_vals, void *stream)
{
	SVT_CHECK(svtgpu_require_device());
	GenParams G;
	SVT_CHECK(fill_params(&G, nrow, nleaf, leaf0, seed, nz_threshold,
			      na_threshold, value_thresholds,
			      n_value_thresholds));
	if (nleaf == 0)
		return SVTGPU_OK;
	const int64_t blocks = (nleaf + 7) / 8;
	SVT_ARG(blocks <= INT32_MAX, "svtgpu_gen_fill: too many leaves");
	cudaStream_t s = (cudaStream_t) stream;
	if (val_type == 0 || d_vals == NULL)
		gen_fill<int32_t, false><<<(unsigned) blocks, 256, 0, s>>>(
			G, d_leaf_ptr, d_offs, (int32_t *) NULL);
	else if (val_type == SVTGPU_DOUBLE)
		gen_fill<double, true><<<(unsigned) blocks, 256, 0, s>>>(
			G, d_leaf_ptr, d_offs, (double *) d_vals);
	else
		gen_fill<int32_t, true><<<(unsigned) blocks, 256, 0, s>>>(
			G, d_leaf_ptr, d_offs, (int32_t *) d_vals);
	SVT_CUDA(cudaGetLastError());
	svtgpu_count_launch(1);
	return SVTGPU_OK;
}
