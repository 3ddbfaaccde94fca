// How long does the first allocation of a multi-GB device array take?
// nvcc -O2 -o alloc_time alloc_time.cu
#include <cstdio>
#include <chrono>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv)
{
	size_t gb = argc > 1 ? atol(argv[1]) : 19;
	int mode = argc > 2 ? atoi(argv[2]) : 0;   // 0 cudaMalloc, 1 cudaMallocAsync
	cudaFree(0);
	cudaMemPool_t pool; cudaDeviceGetDefaultMemPool(&pool, 0);
	unsigned long long keep = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
	for (int rep = 0; rep < 3; rep++) {
		void *p[3];
		double t0 = now();
		for (int k = 0; k < 3; k++) {
			size_t bytes = (k == 0 ? gb / 2 : gb) * (size_t) 1000000000 / (k == 2 ? 40 : 1);
			cudaError_t e = mode ? cudaMallocAsync(&p[k], bytes, 0) : cudaMalloc(&p[k], bytes);
			if (e != cudaSuccess) { printf("alloc failed: %s\n", cudaGetErrorString(e)); return 1; }
		}
		cudaDeviceSynchronize();
		double t1 = now();
		cudaMemsetAsync(p[1], 0, gb * (size_t) 1000000000, 0);
		cudaDeviceSynchronize();
		double t2 = now();
		for (int k = 0; k < 3; k++) { if (mode) cudaFreeAsync(p[k], 0); else cudaFree(p[k]); }
		cudaDeviceSynchronize();
		double t3 = now();
		printf("%s rep %d: alloc %.1f ms, memset(%zu GB) %.1f ms, free %.1f ms\n", mode ? "cudaMallocAsync" : "cudaMalloc", rep, t1 - t0, gb, t2 - t1, t3 - t2);
	}
	return 0;
}
