/* sm_100a PTX helpers: mbarrier, 1-D bulk async copies (TMA), warp reductions. */
#ifndef SVT_PTX_CUH
#define SVT_PTX_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#define SVT_FULL_MASK 0xffffffffu

__device__ __forceinline__ uint32_t svt_smem_u32(const void *p)
{
	return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void svt_mbar_init(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"
		     :: "r"(bar), "r"(count) : "memory");
}

/* make mbarrier initialisation visible to the async proxy (TMA) */
__device__ __forceinline__ void svt_mbar_init_fence(void)
{
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void svt_mbar_arrive(uint32_t bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];"
		     :: "r"(bar) : "memory");
}

__device__ __forceinline__ void svt_mbar_arrive_expect_tx(uint32_t bar,
							   uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
		     :: "r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ uint32_t svt_mbar_try_wait(uint32_t bar,
						      uint32_t parity)
{
	uint32_t ok;
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		"selp.u32 %0, 1, 0, p;\n\t"
		"}"
		: "=r"(ok) : "r"(bar), "r"(parity) : "memory");
	return ok;
}

__device__ __forceinline__ void svt_mbar_wait(uint32_t bar, uint32_t parity)
{
	while (!svt_mbar_try_wait(bar, parity)) { }
}

/* global -> shared bulk copy of `bytes` (multiple of 16; both addresses
   16-byte aligned); completion is signalled on `bar` as transaction bytes. */
__device__ __forceinline__ void svt_bulk_g2s(uint32_t dst, const void *src,
					     uint32_t bytes, uint32_t bar)
{
	asm volatile(
		"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"
		" [%0], [%1], %2, [%3];"
		:: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

/* same with an L2 evict-first policy: data streamed exactly once */
__device__ __forceinline__ uint64_t svt_policy_evict_first(void)
{
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;"
		     : "=l"(pol));
	return pol;
}

__device__ __forceinline__ void svt_bulk_g2s_hint(uint32_t dst,
						  const void *src,
						  uint32_t bytes, uint32_t bar,
						  uint64_t policy)
{
	asm volatile(
		"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"
		".L2::cache_hint [%0], [%1], %2, [%3], %4;"
		:: "r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
		: "memory");
}

__device__ __forceinline__ uint64_t svt_policy_evict_last(void)
{
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;"
		     : "=l"(pol));
	return pol;
}

/* loads / stores that carry an L2 eviction policy */
__device__ __forceinline__ int32_t svt_ldg_hint(const int32_t *p, uint64_t pol)
{
	int32_t r;
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;"
		     : "=r"(r) : "l"(p), "l"(pol));
	return r;
}

__device__ __forceinline__ double svt_ldg_hint(const double *p, uint64_t pol)
{
	double r;
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;"
		     : "=d"(r) : "l"(p), "l"(pol));
	return r;
}

__device__ __forceinline__ void svt_stg_hint(int32_t *p, int32_t v, uint64_t pol)
{
	asm volatile("st.global.L2::cache_hint.s32 [%0], %1, %2;"
		     :: "l"(p), "r"(v), "l"(pol) : "memory");
}

__device__ __forceinline__ void svt_stg_hint(double *p, double v, uint64_t pol)
{
	asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;"
		     :: "l"(p), "d"(v), "l"(pol) : "memory");
}

/* 4- / 8-byte global -> shared asynchronous copies (LDGSTS) and their
   completion groups */
__device__ __forceinline__ void svt_cp_async4(uint32_t dst, const void *src)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;"
		     :: "r"(dst), "l"(src) : "memory");
}

__device__ __forceinline__ void svt_cp_async8(uint32_t dst, const void *src)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;"
		     :: "r"(dst), "l"(src) : "memory");
}

__device__ __forceinline__ void svt_cp_async_commit(void)
{
	asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int N>
__device__ __forceinline__ void svt_cp_async_wait(void)
{
	asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}

/* streaming 16-byte global load that does not allocate in L1 */
__device__ __forceinline__ int4 svt_ldg_stream(const int4 *p)
{
	int4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
		     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}

__device__ __forceinline__ double svt_shfl_xor(double v, int m)
{
	return __shfl_xor_sync(SVT_FULL_MASK, v, m);
}

__device__ __forceinline__ long long svt_shfl_xor(long long v, int m)
{
	return __shfl_xor_sync(SVT_FULL_MASK, v, m);
}

__device__ __forceinline__ double svt_warp_sum(double v)
{
#pragma unroll
	for (int m = 16; m > 0; m >>= 1)
		v += svt_shfl_xor(v, m);
	return v;
}

__device__ __forceinline__ long long svt_warp_sum(long long v)
{
#pragma unroll
	for (int m = 16; m > 0; m >>= 1)
		v += svt_shfl_xor(v, m);
	return v;
}

__device__ __forceinline__ double svt_warp_prod(double v)
{
#pragma unroll
	for (int m = 16; m > 0; m >>= 1)
		v *= svt_shfl_xor(v, m);
	return v;
}

__device__ __forceinline__ double svt_warp_min(double v)
{
#pragma unroll
	for (int m = 16; m > 0; m >>= 1) {
		double o = svt_shfl_xor(v, m);
		v = o < v ? o : v;
	}
	return v;
}

__device__ __forceinline__ double svt_warp_max(double v)
{
#pragma unroll
	for (int m = 16; m > 0; m >>= 1) {
		double o = svt_shfl_xor(v, m);
		v = o > v ? o : v;
	}
	return v;
}

#endif  /* SVT_PTX_CUH */
