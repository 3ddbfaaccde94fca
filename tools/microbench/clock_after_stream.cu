// What is the SM clock during the 3 ms that follow 1.5 ms of HBM streaming at
// full bandwidth (a column reduction), compared with an idle GPU and with a
// shared-memory-heavy kernel before it?  A probe kernel spins on clock64() and
// reads %globaltimer at checkpoints: MHz = d(clock64) / d(globaltimer).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o clock_after_stream clock_after_stream.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void stream_read(const int4 *p, size_t n, int *sink)
{
	int acc = 0;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n;
	     i += (size_t) gridDim.x * blockDim.x) {
		int4 v = p[i];
		acc += v.x + v.y + v.z + v.w;
	}
	if (acc == 0x12345678) *sink = acc;
}

#define NCHK 12
__global__ void probe(long long step, unsigned long long *out)
{
	// one thread per CTA, one CTA per SM
	unsigned long long g0, g;
	long long c0 = clock64(), c;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
	for (int k = 0; k < NCHK; k++) {
		do { c = clock64(); } while (c - c0 < step * (k + 1));
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
		if (blockIdx.x == 0) {
			out[2 * k] = (unsigned long long) (c - c0);
			out[2 * k + 1] = g - g0;
		}
	}
}

static void report(const char *what, unsigned long long *d_out)
{
	unsigned long long h[2 * NCHK];
	cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
	printf("%-28s", what);
	unsigned long long pc = 0, pg = 0;
	for (int k = 0; k < NCHK; k++) {
		double mhz = (double) (h[2 * k] - pc) / (double) (h[2 * k + 1] - pg) * 1e3;
		printf(" %5.0f", mhz);
		pc = h[2 * k]; pg = h[2 * k + 1];
	}
	printf("  MHz per 0.25 ms\n");
}

int main()
{
	size_t bytes = (size_t) 9400 << 20;
	int4 *buf; int *sink; unsigned long long *d_out;
	cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4); cudaMalloc(&d_out, 16 * NCHK);
	cudaMemset(buf, 1, bytes);
	const long long step = 491250;  // 0.25 ms at 1965 MHz
	for (int rep = 0; rep < 3; rep++) {
		cudaDeviceSynchronize();
		probe<<<148, 1>>>(step, d_out);
		cudaDeviceSynchronize();
		report("idle -> probe", d_out);
		for (int k = 0; k < 20; k++) {
			stream_read<<<148 * 8, 256>>>(buf, bytes / 16, sink);
		}
		probe<<<148, 1>>>(step, d_out);
		cudaDeviceSynchronize();
		report("20 x stream 9.4 GB -> probe", d_out);
		stream_read<<<148 * 8, 256>>>(buf, bytes / 16, sink);
		probe<<<148, 1>>>(step, d_out);
		cudaDeviceSynchronize();
		report("1 x stream -> probe", d_out);
	}
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	cudaEventRecord(a);
	stream_read<<<148 * 8, 256>>>(buf, bytes / 16, sink);
	cudaEventRecord(b); cudaDeviceSynchronize();
	float ms; cudaEventElapsedTime(&ms, a, b);
	printf("stream_read: %.3f ms = %.0f GB/s\n", ms, bytes / ms / 1e6);
	return 0;
}
