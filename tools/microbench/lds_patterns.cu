// Shared-memory pipe cost (SM cycles per warp instruction, 16 warps/SM) of the
// access patterns the crossprod gather can use.  Addresses are computed before
// the timed loop; the loop is UNROLL loads + one add each.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_patterns lds_patterns.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4000
#define UNROLL 8
#define PITCH 400
#define NROWS 520

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int PAT>
__global__ void __launch_bounds__(512, 1) k(unsigned long long *out, double *sink)
{
	extern __shared__ __align__(16) unsigned char sm[];
	for (int i = threadIdx.x; i < (PITCH * NROWS + 1024) / 4; i += blockDim.x)
		((uint32_t *) sm)[i] = i & 1023;
	__syncthreads();
	const int lane = threadIdx.x & 31, half = lane >> 4, c = lane & 15;
	const uint32_t base = (uint32_t) __cvta_generic_to_shared(sm);
	uint32_t seed = (threadIdx.x >> 5) * 2654435761u + blockIdx.x * 97u + 12345u;
	uint32_t addr[UNROLL];
	int src[UNROLL];
#pragma unroll
	for (int u = 0; u < UNROLL; u++) {
		// rows: one per half-warp, per quarter-warp, per lane
		uint32_t rw = lcg(seed);
		uint32_t r_half = __shfl_sync(0xffffffffu, rw * (lane + 1), half * 16) % NROWS;
		uint32_t r_quarter = __shfl_sync(0xffffffffu, rw * (lane + 1), (lane >> 3) * 8) % NROWS;
		uint32_t r_lane = (rw * (lane + 1) * 2246822519u >> 7) % NROWS;
		uint32_t rec = (rw >> 3) & 31;
		src[u] = rec;
		switch (PAT) {
		case 0: addr[u] = base + r_half * PITCH + c * 16; break;             // LDS.128 256 B per half
		case 1: addr[u] = base + r_half * PITCH + 256 + c * 8; break;        // LDS.64 128 B per half
		case 2: addr[u] = base + r_half * PITCH + 384 + (c & 1) * 8; break;  // LDS.64 2 addrs per half
		case 3: addr[u] = base + (rec * 2 + half) * 16; break;              // LDS.128 broadcast per half
		case 5: addr[u] = base + (rec * 2 + half) * 8; break;               // LDS.64 broadcast per half
		case 6: addr[u] = base + (rec * 2 + half) * 4; break;               // LDS.32 broadcast per half
		case 7: addr[u] = base + r_half * PITCH + (c & 7) * 16; break;       // LDS.128 lanes 8-15 duplicate 0-7
		case 9: addr[u] = base + r_lane * PITCH + 384; break;                // LDS.128 32 random rows
		case 11: addr[u] = base + r_quarter * PITCH + (lane & 7) * 16; break;// LDS.128 quarter-warp rows
		case 12: addr[u] = base + (rec * 4 + (lane >> 3)) * 16; break;       // LDS.128 broadcast per quarter
		case 13: addr[u] = base + r_half * PITCH + 256 + (c < 9 ? c : 0) * 16; break; // LDS.128 9 lanes + dup
		default: addr[u] = base; break;
		}
	}
	uint32_t acci = seed;
	__syncthreads();
	long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < ITERS; it++) {
#pragma unroll
		for (int u = 0; u < UNROLL; u++) {
			if (PAT == 0 || PAT == 7 || PAT == 9 || PAT == 11 || PAT == 13 || PAT == 3 || PAT == 12) {
				uint32_t a, b, d, e;
				asm volatile("ld.volatile.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(d), "=r"(e) : "r"(addr[u]) : "memory");
				acci += a;
			} else if (PAT == 1 || PAT == 2 || PAT == 5) {
				uint32_t a, b;
				asm volatile("ld.volatile.shared.v2.b32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr[u]) : "memory");
				acci += a;
			} else if (PAT == 6) {
				uint32_t a;
				asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(a) : "r"(addr[u]) : "memory");
				acci += a;
			} else if (PAT == 4) {
				acci ^= __shfl_sync(0xffffffffu, seed, src[u]);
			} else if (PAT == 14) {
				acci ^= __shfl_sync(0xffffffffu, seed, src[u]);
				acci ^= __shfl_sync(0xffffffffu, seed + 1, src[u]);
			}
		}
	}
	long long t1 = clock64();
	sink[blockIdx.x * blockDim.x + threadIdx.x] = (double) acci;
	if (threadIdx.x == 0)
		out[blockIdx.x] = (unsigned long long) (t1 - t0);
}

template <int PAT> void run(const char *name, unsigned long long *d, double *sink)
{
	size_t smem = (size_t) PITCH * NROWS + 1024;
	cudaFuncSetAttribute(k<PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	k<PAT><<<148, 512, smem>>>(d, sink);
	cudaDeviceSynchronize();
	unsigned long long h[148];
	cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
	double cyc = 0;
	for (int i = 0; i < 148; i++) cyc += (double) h[i];
	cyc /= 148;
	printf("%-56s %6.2f SM-cycles per warp-instruction (%s)\n", name, cyc / (16.0 * ITERS * UNROLL), cudaGetErrorString(cudaGetLastError()));
}

int main()
{
	unsigned long long *d;
	double *sink;
	cudaMalloc(&d, 8 * 148);
	cudaMalloc(&sink, 8 * 148 * 512);
	run<0>("LDS.128 16 lanes x 16 B per half, 2 rows (512 B)", d, sink);
	run<11>("LDS.128 8 lanes x 16 B per quarter, 4 rows (512 B)", d, sink);
	run<7>("LDS.128 128 B per half (lanes 8-15 duplicate 0-7)", d, sink);
	run<13>("LDS.128 144 B per half (9 lanes, rest duplicate lane 0)", d, sink);
	run<1>("LDS.64  16 lanes x 8 B per half, 2 rows (256 B)", d, sink);
	run<2>("LDS.64  2 addresses per half (tail, 32 B)", d, sink);
	run<3>("LDS.128 broadcast: 1 record per half (32 B)", d, sink);
	run<12>("LDS.128 broadcast: 1 record per quarter (64 B)", d, sink);
	run<5>("LDS.64  broadcast: 1 value per half", d, sink);
	run<6>("LDS.32  broadcast: 1 word per half", d, sink);
	run<4>("SHFL.IDX 32-bit (independent)", d, sink);
	run<14>("SHFL.IDX 2 x 32-bit (independent)", d, sink);
	run<9>("LDS.128 32 random rows, 16 B each (tail gather)", d, sink);
	return 0;
}
