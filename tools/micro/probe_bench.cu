// tools/micro/probe_bench.cu -- what does a random 16/32/64-byte probe of an HBM-resident
// table cost on this chip?  Prints probes/s and useful GB/s for several access widths,
// L2 fetch-granularity limits and loads in flight per thread.  (Measurement tool only.)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

template <int BYTES, int MLP>
__global__ void probe(const uint8_t* __restrict__ t, uint64_t nslots, uint64_t n_probes, uint64_t* sink)
{
	uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	uint64_t acc = 0;
	for (uint64_t i = tid * MLP; i < n_probes; i += stride * MLP) {
		uint64_t v[MLP][4];
#pragma unroll
		for (int m = 0; m < MLP; ++m) {
			uint64_t s = __umul64hi(mix(i + m + 12345), nslots);
			const uint8_t* p = t + s * BYTES;
			if (BYTES == 16) {
				asm("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(v[m][0]), "=l"(v[m][1]) : "l"(p));
				v[m][2] = v[m][3] = 0;
			} else if (BYTES == 32) {
				asm("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[m][0]), "=l"(v[m][1]), "=l"(v[m][2]), "=l"(v[m][3]) : "l"(p));
			} else { // 64: two 32-byte loads of one 64-byte bucket
				uint64_t a, b, c, d;
				asm("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[m][0]), "=l"(v[m][1]), "=l"(v[m][2]), "=l"(v[m][3]) : "l"(p));
				asm("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p + 32));
				v[m][0] ^= a ^ b ^ c ^ d;
			}
		}
#pragma unroll
		for (int m = 0; m < MLP; ++m)
			acc += v[m][0] ^ v[m][1] ^ v[m][2] ^ v[m][3];
	}
	if (acc == 0x1234567)
		*sink = acc;
}

template <int BYTES, int MLP>
void run(const uint8_t* t, uint64_t bytes, uint64_t n_probes, uint64_t* sink, int blocks_per_sm, int threads)
{
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	int grid = 148 * blocks_per_sm;
	probe<BYTES, MLP><<<grid, threads>>>(t, bytes / BYTES, n_probes / 8, sink);
	cudaEventRecord(e0);
	probe<BYTES, MLP><<<grid, threads>>>(t, bytes / BYTES, n_probes, sink);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	printf("  bytes=%2d mlp=%d blocks/sm=%d thr=%d : %7.2f ms  %6.1f Gprobe/s  %7.1f GB/s useful\n", BYTES, MLP, blocks_per_sm, threads, ms,
	    n_probes / ms / 1e6, n_probes * (double)BYTES / ms / 1e6);
}

int main()
{
	uint64_t bytes = 3ull << 30;
	uint8_t* t;
	uint64_t* sink;
	cudaMalloc(&t, bytes);
	cudaMalloc(&sink, 8);
	cudaMemset(t, 1, bytes);
	uint64_t n = 1ull << 29;
	for (int gran : {0, 32, 64, 128}) {
		if (gran) {
			cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
			size_t g = 0;
			cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
			printf("L2 fetch granularity limit set to %d (%s) -> reads back %zu\n", gran, cudaGetErrorString(e), g);
		} else {
			size_t g = 0;
			cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
			printf("default L2 fetch granularity limit = %zu\n", g);
		}
		run<16, 4>(t, bytes, n, sink, 8, 256);
		run<32, 1>(t, bytes, n, sink, 8, 256);
		run<32, 4>(t, bytes, n, sink, 8, 256);
		run<32, 8>(t, bytes, n, sink, 8, 256);
		run<32, 4>(t, bytes, n, sink, 2, 256);
		run<64, 4>(t, bytes, n, sink, 8, 256);
	}
	return 0;
}
