// Read-only streaming bandwidth by load width and loads in flight: is the
// gap between the column kernels (6.2 TB/s, 4-byte loads) and a plain int4
// loop (7.0 TB/s) the load width?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_width stream_width.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <typename V, int U>
__global__ void __launch_bounds__(256) rd(const V *p, size_t n, int *sink)
{
	int acc = 0;
	const size_t stride = (size_t) gridDim.x * blockDim.x;
	size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	for (; i + (U - 1) * stride < n; i += U * stride) {
		V v[U];
#pragma unroll
		for (int k = 0; k < U; k++) v[k] = p[i + k * stride];
#pragma unroll
		for (int k = 0; k < U; k++) {
			const int *w = (const int *) &v[k];
#pragma unroll
			for (int j = 0; j < (int) (sizeof(V) / 4); j++) acc += w[j];
		}
	}
	if (acc == 0x12345678) *sink = acc;
}

template <typename V, int U> void run(const char *name, const void *buf, size_t bytes, int blocks_per_sm, int *sink)
{
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	float best = 1e9f;
	for (int rep = 0; rep < 5; rep++) {
		cudaEventRecord(a);
		rd<V, U><<<148 * blocks_per_sm, 256>>>((const V *) buf, bytes / sizeof(V), sink);
		cudaEventRecord(b); cudaDeviceSynchronize();
		float ms; cudaEventElapsedTime(&ms, a, b);
		if (ms < best) best = ms;
	}
	printf("%-28s %d blocks/SM  %.3f ms  %.0f GB/s\n", name, blocks_per_sm, best, bytes / best / 1e6);
}

int main()
{
	size_t bytes = (size_t) 9400 << 20;
	void *buf; int *sink;
	cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
	cudaMemset(buf, 1, bytes);
	run<int, 1>("int x1", buf, bytes, 8, sink);
	run<int, 4>("int x4", buf, bytes, 8, sink);
	run<int, 8>("int x8", buf, bytes, 8, sink);
	run<int, 16>("int x16", buf, bytes, 8, sink);
	run<int2, 8>("int2 x8", buf, bytes, 8, sink);
	run<int4, 1>("int4 x1", buf, bytes, 8, sink);
	run<int4, 2>("int4 x2", buf, bytes, 8, sink);
	run<int4, 4>("int4 x4", buf, bytes, 8, sink);
	run<int4, 4>("int4 x4", buf, bytes, 4, sink);
	run<int, 8>("int x8", buf, bytes, 4, sink);
	return 0;
}
