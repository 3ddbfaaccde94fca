// Shared-memory atomic / read-modify-write throughput per SM for the row
// histogram's access patterns: lanes per clock and SM, 1024 threads per SM,
// 200 KB of 32-bit cells.  Addresses are computed before the timed loop.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atoms_patterns atoms_patterns.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define NCELL 50000
#define U 16
#define ITER 256

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
	x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
	return x;
}

// mode 0: atomicAdd, random cells; 1: atomicAdd, conflict-free (lane = bank);
// 2: plain ld + st, random cells; 3: plain, conflict-free;
// 4: atomicAdd, random cells, sorted ascending across the warp's lanes (as
//    the offsets of a leaf are); 5: atomicOr random
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(long long *cycles, uint32_t *sink)
{
	extern __shared__ uint32_t cell[];
	for (int i = threadIdx.x; i < NCELL; i += blockDim.x) cell[i] = 0;
	const int lane = threadIdx.x & 31;
	uint32_t a[U];
#pragma unroll
	for (int u = 0; u < U; u++) {
		uint32_t h = hash32(threadIdx.x * 131u + u * 7919u + blockIdx.x * 104729u);
		uint32_t c;
		if (MODE == 1 || MODE == 3) c = ((h % (NCELL / 32)) * 32 + lane);
		else if (MODE == 4) c = (uint32_t) (((hash32((threadIdx.x >> 5) * 977u + u) % 16) * 32 + lane) * 97u + (h % 97u)) % NCELL;
		else c = h % NCELL;
		a[u] = c * 4u;
	}
	__syncthreads();
	const uint32_t base = (uint32_t) __cvta_generic_to_shared(cell);
	long long t0 = clock64();
	for (int it = 0; it < ITER; it++) {
#pragma unroll
		for (int u = 0; u < U; u++) {
			const uint32_t ad = base + a[u];
			if (MODE == 0 || MODE == 1 || MODE == 4) {
				asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(ad), "r"(1u << ((it & 1) * 16)) : "memory");
			} else if (MODE == 5) {
				asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(ad), "r"(1u << (it & 31)) : "memory");
			} else {
				uint32_t v;
				asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ad) : "memory");
				asm volatile("st.shared.u32 [%0], %1;" :: "r"(ad), "r"(v + 1u) : "memory");
			}
		}
	}
	__syncthreads();
	long long t1 = clock64();
	if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
	uint32_t s = 0;
	for (int i = threadIdx.x; i < NCELL; i += blockDim.x) s += cell[i];
	if (s == 0xdeadbeef) sink[0] = s;
}

template <int MODE> void run(const char *name)
{
	long long *d, h[148]; uint32_t *sink;
	cudaMalloc(&d, 148 * 8); cudaMalloc(&sink, 4);
	cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, NCELL * 4);
	for (int rep = 0; rep < 2; rep++) k<MODE><<<148, 1024, NCELL * 4>>>(d, sink);
	cudaDeviceSynchronize();
	cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
	double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
	double lanes = 1024.0 * U * ITER;
	printf("%-44s %8.0f cycles  %.2f lanes/clk/SM  (%.2f clk per warp instruction)\n", name, c, lanes / c, c / (lanes / 32));
	cudaFree(d); cudaFree(sink);
}

int main()
{
	run<0>("red.shared.add random cells");
	run<1>("red.shared.add lane = bank");
	run<4>("red.shared.add ascending within the warp");
	run<5>("red.shared.or random cells");
	run<2>("ld + st random cells");
	run<3>("ld + st lane = bank");
	return 0;
}
