// arks_links.cuh -- kernel 3: per-barcode tallies -> pairwise link counters.
//
// Replaces pairContigs (Arcs/Arcs.cpp:1378-1435) with headOrTail / normalEstimation
// (:833-861) folded into an exact host-built decision table: for a given
// (min_reads, error_percent) the predicate depends only on (max, sum) and is monotone in
// max, so min_max[sum] = smallest passing max decides it with integer compares
// (the float/double/erf expression itself is evaluated on the host, in arks_api.cu).
//
// Pipeline over the device imap (open-address table of {barcode<<32|contig, head, tail}):
//   imap_maxsum_kernel    largest head+tail (sizes the decision table)
//   imap_count_kernel     per barcode: number of contigs whose head/tail test passes and
//                         whose barcode multiplicity is inside [min_mult, max_mult]
//   scan_*                exclusive prefix sum over barcodes (hand-written, 3 phases)
//   imap_scatter_kernel   counting-sort the passing rows by barcode
//   pair_kernel           one warp per barcode: all contig pairs, ordered by the
//                         host-supplied std::string rank, orientation index
//                         2*(!Ahead)+(!Bhead) (:1418-1428), accumulated into the pmap
//                         hash {a<<32|b -> counts[4]}
#pragma once
#include "arks_device.cuh"
#include "arks_map.cuh"

namespace arks {

struct LinkParams
{
	const unsigned long long* imap;
	uint64_t imap_cap;
	const int32_t* mult;
	uint32_t n_barcodes;
	int min_mult, max_mult;
	const uint32_t* min_max; // decision table
	uint32_t n_min_max;
};

// slot tables: word 0 of every slot = empty key (all ones), the other words = 0
__global__ void init_slots_kernel(unsigned long long* t, uint64_t n_words, uint32_t words_per_slot)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_words; i += (uint64_t)gridDim.x * blockDim.x)
		t[i] = (i % words_per_slot == 0) ? kEmptyKey : 0ull;
}

__global__ void imap_maxsum_kernel(const unsigned long long* imap, uint64_t cap, uint32_t* maxsum)
{
	uint32_t m = 0;
	for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < cap; s += (uint64_t)gridDim.x * blockDim.x) {
		if (imap[2 * s] == kEmptyKey)
			continue;
		unsigned long long ht = imap[2 * s + 1];
		m = max(m, (uint32_t)ht + (uint32_t)(ht >> 32));
	}
	m = __reduce_max_sync(0xFFFFFFFFu, m);
	if ((threadIdx.x & 31) == 0 && m)
		atomicMax(maxsum, m);
}

// headOrTail through the decision table: returns 0 (invalid), 1 (valid, tail), 3 (valid, head)
__device__ __forceinline__ uint32_t head_or_tail(const LinkParams& L, uint32_t head, uint32_t tail)
{
	uint32_t sum = head + tail, mx = max(head, tail);
	if (sum >= L.n_min_max)
		return 0; // cannot happen: table covers the largest sum present
	if (mx < L.min_max[sum])
		return 0;
	return mx == head ? 3u : 1u;
}

__device__ __forceinline__ bool row_passes(const LinkParams& L, unsigned long long key, unsigned long long ht, uint32_t& barcode, uint32_t& packed)
{
	barcode = (uint32_t)(key >> 32);
	if (barcode >= L.n_barcodes)
		return false;
	int mu = L.mult[barcode];
	if (mu < L.min_mult || mu > L.max_mult)
		return false;
	uint32_t d = head_or_tail(L, (uint32_t)ht, (uint32_t)(ht >> 32));
	if (!d)
		return false;
	packed = (uint32_t)key | ((d & 2u) ? 0x80000000u : 0u);
	return true;
}

__global__ void imap_count_kernel(LinkParams L, uint32_t* cnt)
{
	for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < L.imap_cap; s += (uint64_t)gridDim.x * blockDim.x) {
		unsigned long long key = L.imap[2 * s];
		if (key == kEmptyKey)
			continue;
		uint32_t b, packed;
		if (row_passes(L, key, L.imap[2 * s + 1], b, packed))
			atomicAdd(&cnt[b], 1u);
	}
}

__global__ void imap_scatter_kernel(LinkParams L, const uint32_t* offs, uint32_t* fill, uint32_t* rows)
{
	for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < L.imap_cap; s += (uint64_t)gridDim.x * blockDim.x) {
		unsigned long long key = L.imap[2 * s];
		if (key == kEmptyKey)
			continue;
		uint32_t b, packed;
		if (row_passes(L, key, L.imap[2 * s + 1], b, packed))
			rows[offs[b] + atomicAdd(&fill[b], 1u)] = packed;
	}
}

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

// out[i] = exclusive prefix; out[n] = grand total
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

__global__ void pair_count_kernel(const uint32_t* cnt, uint32_t n_barcodes, unsigned long long* events)
{
	unsigned long long e = 0;
	for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < n_barcodes; b += gridDim.x * blockDim.x) {
		unsigned long long v = cnt[b];
		e += v * (v - 1) / 2;
	}
	for (int o = 16; o > 0; o >>= 1)
		e += __shfl_down_sync(0xFFFFFFFFu, e, o);
	if ((threadIdx.x & 31) == 0 && e)
		atomicAdd(events, e);
}

// pmap slot: 32 B = {u64 key = a<<32|b, u32 counts[4], u64 pad}
__device__ __forceinline__ void pmap_add(unsigned long long* pmap, uint64_t mask, unsigned long long* count, uint32_t a, uint32_t b, uint32_t orient)
{
	unsigned long long key = ((unsigned long long)a << 32) | b;
	uint64_t slot = mix64(key) & mask;
	while (true) {
		unsigned long long* p = pmap + 4 * slot;
		unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(p);
		if (cur == kEmptyKey) {
			cur = atomicCAS(p, (unsigned long long)kEmptyKey, key);
			if (cur == kEmptyKey) {
				atomicAdd(count, 1ull);
				cur = key;
			}
		}
		if (cur == key) {
			atomicAdd(reinterpret_cast<uint32_t*>(p + 1) + orient, 1u);
			return;
		}
		slot = (slot + 1) & mask;
	}
}

__global__ void pair_kernel(const uint32_t* offs, const uint32_t* rows, uint32_t n_barcodes, const uint32_t* rank,
    unsigned long long* pmap, uint64_t pmap_mask, unsigned long long* pmap_count)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t bc = gwarp; bc < n_barcodes; bc += nwarps) {
		const uint32_t r0 = offs[bc], r1 = offs[bc + 1];
		for (uint32_t i = r0; i + 1 < r1; ++i) {
			const uint32_t ri = rows[i];
			const uint32_t ci = ri & 0x7FFFFFFFu, hi = ri >> 31;
			const uint32_t rki = rank[ci];
			for (uint32_t j = i + 1 + lane; j < r1; j += 32) {
				const uint32_t rj = rows[j];
				const uint32_t cj = rj & 0x7FFFFFFFu, hj = rj >> 31;
				const uint32_t rkj = rank[cj];
				if (rki == rkj)
					continue;
				const bool i_first = rki < rkj;
				const uint32_t a = i_first ? ci : cj, b = i_first ? cj : ci;
				const uint32_t ah = i_first ? hi : hj, bh = i_first ? hj : hi;
				pmap_add(pmap, pmap_mask, pmap_count, a, b, (ah ? 0u : 2u) + (bh ? 0u : 1u));
			}
		}
	}
}

__global__ void pmap_export_kernel(const unsigned long long* pmap, uint64_t cap, uint32_t* a, uint32_t* b, uint32_t* counts,
    unsigned long long* counter, uint64_t out_cap)
{
	for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < cap; s += (uint64_t)gridDim.x * blockDim.x) {
		unsigned long long key = pmap[4 * s];
		if (key == kEmptyKey)
			continue;
		unsigned long long i = atomicAdd(counter, 1ull);
		if (i < out_cap) {
			a[i] = (uint32_t)(key >> 32);
			b[i] = (uint32_t)key;
			const uint32_t* c = reinterpret_cast<const uint32_t*>(pmap + 4 * s + 1);
			counts[4 * i + 0] = c[0];
			counts[4 * i + 1] = c[1];
			counts[4 * i + 2] = c[2];
			counts[4 * i + 3] = c[3];
		}
	}
}

__global__ void imap_export_kernel(const unsigned long long* imap, uint64_t cap, uint32_t* barcode, uint32_t* contig,
    uint32_t* head, uint32_t* tail, unsigned long long* counter, uint64_t out_cap)
{
	for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < cap; s += (uint64_t)gridDim.x * blockDim.x) {
		unsigned long long key = imap[2 * s];
		if (key == kEmptyKey)
			continue;
		unsigned long long i = atomicAdd(counter, 1ull);
		if (i < out_cap) {
			unsigned long long ht = imap[2 * s + 1];
			barcode[i] = (uint32_t)(key >> 32);
			contig[i] = (uint32_t)key;
			head[i] = (uint32_t)ht;
			tail[i] = (uint32_t)(ht >> 32);
		}
	}
}

// re-inserts every row of an old imap table into a new (larger) one
__global__ void imap_rehash_kernel(const unsigned long long* old_imap, uint64_t old_cap, unsigned long long* imap, uint64_t mask,
    unsigned long long* count)
{
	for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < old_cap; s += (uint64_t)gridDim.x * blockDim.x) {
		unsigned long long key = old_imap[2 * s];
		if (key == kEmptyKey)
			continue;
		unsigned long long ht = old_imap[2 * s + 1];
		imap_add(imap, mask, count, (uint32_t)(key >> 32), (uint32_t)key, (uint32_t)ht, (uint32_t)(ht >> 32));
	}
}

__global__ void imap_add_rows_kernel(const uint32_t* barcode, const uint32_t* contig, const uint32_t* head, const uint32_t* tail,
    uint64_t n, unsigned long long* imap, uint64_t mask, unsigned long long* count)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		imap_add(imap, mask, count, barcode[i], contig[i], head[i], tail[i]);
}

} // namespace arks
