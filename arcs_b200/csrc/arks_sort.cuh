// arks_sort.cuh -- device-side ordering primitives for the pair-link map (sm_100a, hand-written):
//
//  * exclusive scan of uint32 (three phases, n up to 2^32-1);
//  * stable LSD radix sort of (uint64 key, uint32 value) records, 8 bits per pass: per-tile digit
//    histograms -> one scan over the digit-major histogram matrix -> stable scatter (every warp owns 512
//    consecutive records of its tile; ranks inside a warp come from match.any, across the warps of a tile
//    from a per-digit prefix in shared memory);
//  * run heads / compaction (unique) and a lower-bound scatter, used by the multi-GPU merge.
//
// The reference keeps its pair-link map in a std::map<pair<string,string>, vector<unsigned>> (Arcs/Arcs.h:115)
// whose iteration order -- (name a, name b) under std::string '<' -- defines the order of edges and vertices
// of _original.gv (createGraph, Arcs/Arcs.cpp:1475-1526).  Keys here are (rank a << 32 | rank b) with the
// host-supplied std::string ranks, so an integer sort reproduces that order.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace arks {

// ---- exclusive scan of uint32 (n up to 2^32-1), three phases -----------------------------
constexpr int kScanBlock = 1024;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total)
{
	__shared__ uint32_t warp_sums[32];
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inc = v;
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
		if (lane >= (uint32_t)o)
			inc += t;
	}
	if (lane == 31)
		warp_sums[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		uint32_t ws = warp_sums[lane];
		uint32_t winc = ws;
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
			if (lane >= (uint32_t)o)
				winc += t;
		}
		warp_sums[lane] = winc - ws; // exclusive
		if (lane == 31)
			*total = winc;
	}
	__syncthreads();
	uint32_t r = inc - v + warp_sums[warp];
	__syncthreads();
	return r;
}

__global__ void __launch_bounds__(kScanBlock) scan_block_sums_kernel(const uint32_t* in, uint64_t n, uint32_t* block_sums)
{
	__shared__ uint32_t total;
	uint64_t i = blockIdx.x * (uint64_t)kScanBlock + threadIdx.x;
	block_exclusive_scan(i < n ? in[i] : 0u, &total);
	if (threadIdx.x == 0)
		block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of block_sums in place
__global__ void __launch_bounds__(kScanBlock) scan_sums_kernel(uint32_t* block_sums, uint32_t n_blocks)
{
	__shared__ uint32_t total;
	uint32_t carry = 0;
	for (uint32_t base = 0; base < n_blocks; base += kScanBlock) {
		uint32_t i = base + threadIdx.x;
		uint32_t v = i < n_blocks ? block_sums[i] : 0u;
		uint32_t ex = block_exclusive_scan(v, &total);
		if (i < n_blocks)
			block_sums[i] = ex + carry;
		carry += total;
		__syncthreads();
	}
}

// out[i] = exclusive prefix; out[n] = grand total (out may alias in)
__global__ void __launch_bounds__(kScanBlock)
scan_apply_kernel(const uint32_t* in, uint64_t n, const uint32_t* block_sums, uint32_t* out)
{
	__shared__ uint32_t total;
	uint64_t i = blockIdx.x * (uint64_t)kScanBlock + threadIdx.x;
	uint32_t v = i < n ? in[i] : 0u;
	uint32_t ex = block_exclusive_scan(v, &total) + block_sums[blockIdx.x];
	if (i < n)
		out[i] = ex;
	if (i == n - 1)
		out[n] = ex + v;
}

// ---- LSD radix sort, 8 bits per pass --------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 16;                            // records per thread
constexpr int kSortTile = kSortThreads * kSortItems;      // 4096 records per CTA
constexpr int kSortWarpSpan = 32 * kSortItems;            // 512 consecutive records per warp
constexpr int kRadix = 256;

// hist[digit * n_tiles + tile] = number of records of the tile with that digit
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const unsigned long long* __restrict__ keys, uint64_t n, uint32_t shift, uint32_t* __restrict__ hist, uint32_t n_tiles)
{
	__shared__ uint32_t h[kRadix];
	h[threadIdx.x] = 0;
	__syncthreads();
	const uint64_t base = blockIdx.x * (uint64_t)kSortTile;
#pragma unroll 4
	for (int i = 0; i < kSortItems; ++i) {
		const uint64_t idx = base + (uint64_t)i * kSortThreads + threadIdx.x;
		if (idx < n)
			atomicAdd(&h[(uint32_t)(keys[idx] >> shift) & 255u], 1u);
	}
	__syncthreads();
	hist[threadIdx.x * (uint64_t)n_tiles + blockIdx.x] = h[threadIdx.x];
}

// offs = exclusive scan of hist (same layout).  Stable: records with equal digits keep their order.
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const unsigned long long* __restrict__ keys_in, const uint32_t* __restrict__ val_in,
    unsigned long long* __restrict__ keys_out, uint32_t* __restrict__ val_out, uint64_t n, uint32_t shift,
    const uint32_t* __restrict__ offs, uint32_t n_tiles)
{
	__shared__ uint32_t cnt[kSortWarps][kRadix];
	__shared__ uint32_t gbase[kRadix];
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int w = 0; w < kSortWarps; ++w)
		cnt[w][threadIdx.x] = 0;
	gbase[threadIdx.x] = offs[threadIdx.x * (uint64_t)n_tiles + blockIdx.x];
	__syncthreads();
	const uint64_t base = blockIdx.x * (uint64_t)kSortTile + (uint64_t)warp * kSortWarpSpan;
	unsigned long long k[kSortItems];
	uint32_t rank[kSortItems];
	const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
	for (int r = 0; r < kSortItems; ++r) {
		const uint64_t idx = base + (uint64_t)r * 32 + lane;
		const bool valid = idx < n;
		k[r] = valid ? keys_in[idx] : ~0ull;
		const uint32_t digit = (uint32_t)(k[r] >> shift) & 255u;
		// lanes without a record form groups of their own (digit codes >= 256) and are never counted
		const uint32_t peers = __match_any_sync(0xFFFFFFFFu, valid ? digit : 256u + lane);
		const uint32_t prev = valid ? cnt[warp][digit] : 0u;
		rank[r] = prev + __popc(peers & lt);
		__syncwarp();
		if (valid && (peers & lt) == 0u) // lowest lane of the group
			cnt[warp][digit] = prev + __popc(peers);
		__syncwarp();
	}
	__syncthreads();
	{ // per digit: exclusive prefix over the warps of the tile
		uint32_t run = 0;
		for (int w = 0; w < kSortWarps; ++w) {
			const uint32_t t = cnt[w][threadIdx.x];
			cnt[w][threadIdx.x] = run;
			run += t;
		}
	}
	__syncthreads();
#pragma unroll
	for (int r = 0; r < kSortItems; ++r) {
		const uint64_t idx = base + (uint64_t)r * 32 + lane;
		if (idx < n) {
			const uint32_t digit = (uint32_t)(k[r] >> shift) & 255u;
			const uint64_t pos = (uint64_t)gbase[digit] + cnt[warp][digit] + rank[r];
			keys_out[pos] = k[r];
			val_out[pos] = val_in[idx];
		}
	}
}

__global__ void iota_kernel(uint32_t* v, uint64_t n)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		v[i] = (uint32_t)i;
}

// ---- sorted keys -> run heads -> unique keys -----------------------------------------------------
__global__ void run_heads_kernel(const unsigned long long* __restrict__ keys, uint64_t n, uint32_t* __restrict__ head)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// pos = exclusive scan of the run-head flags (pos[n] = number of runs)
__global__ void compact_unique_kernel(const unsigned long long* __restrict__ keys, uint64_t n, const uint32_t* __restrict__ pos,
    unsigned long long* __restrict__ out)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		if (pos[i + 1] != pos[i])
			out[pos[i]] = keys[i];
}

// dense[4 * lower_bound(uni, key[i]) ..] = counts[4 i ..] for every local key (all of them are in uni)
__global__ void scatter_counts_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ counts, uint64_t n,
    const unsigned long long* __restrict__ uni, uint64_t n_uni, uint32_t* __restrict__ dense)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		const unsigned long long key = keys[i];
		uint64_t lo = 0, hi = n_uni;
		while (lo < hi) {
			const uint64_t mid = (lo + hi) >> 1;
			if (uni[mid] < key)
				lo = mid + 1;
			else
				hi = mid;
		}
		const uint4 c = reinterpret_cast<const uint4*>(counts)[i];
		reinterpret_cast<uint4*>(dense)[lo] = c;
	}
}

} // namespace arks
