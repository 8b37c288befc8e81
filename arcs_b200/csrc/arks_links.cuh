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
//   scan_*                exclusive prefix sum over barcodes (arks_sort.cuh)
//   imap_scatter_kernel   counting-sort the passing rows by barcode
//   pair_kernel           one thread per row: the row against the later rows of its barcode,
//                         contigs ordered by the host-supplied std::string rank, orientation
//                         index 2*(!Ahead)+(!Bhead) (:1418-1428), accumulated into the pmap
//                         hash {rank a << 32 | rank b -> counts[4]}
//   pmap_collect_kernel   occupied slots -> (key, slot) records
//   radix sort            (arks_sort.cuh) by key = std::map<pair<string,string>> iteration order
//   pmap_gather_counts_kernel / pmap_names_kernel   sorted rows, ready for one device->host copy
#pragma once
#include "arks_device.cuh"
#include "arks_map.cuh"
#include "arks_sort.cuh"

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

__global__ void imap_scatter_kernel(LinkParams L, const uint32_t* offs, uint32_t* fill, uint32_t* rows, uint32_t* row_bc)
{
	for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < L.imap_cap; s += (uint64_t)gridDim.x * blockDim.x) {
		unsigned long long key = L.imap[2 * s];
		if (key == kEmptyKey)
			continue;
		uint32_t b, packed;
		if (row_passes(L, key, L.imap[2 * s + 1], b, packed))
		{
			const uint32_t at = offs[b] + atomicAdd(&fill[b], 1u);
			rows[at] = packed;
			row_bc[at] = b;
		}
	}
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

// pmap slot: 32 B = {u64 key = rank a << 32 | rank b, u32 counts[4], u64 pad}.  Returns false when the table has
// reached its fill limit (the host then doubles it and runs the pass again).
__device__ __forceinline__ bool pmap_add(unsigned long long* pmap, uint64_t mask, unsigned long long* count, uint64_t limit,
    unsigned long long key, uint32_t orient)
{
	uint64_t slot = mix64(key) & mask;
	while (true) {
		unsigned long long* p = pmap + 4 * slot;
		unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(p);
		if (cur == kEmptyKey) {
			if (*reinterpret_cast<volatile unsigned long long*>(count) >= limit)
				return false;
			cur = atomicCAS(p, (unsigned long long)kEmptyKey, key);
			if (cur == kEmptyKey) {
				atomicAdd(count, 1ull);
				cur = key;
			}
		}
		if (cur == key) {
			atomicAdd(reinterpret_cast<uint32_t*>(p + 1) + orient, 1u);
			return true;
		}
		slot = (slot + 1) & mask;
	}
}

// rows sorted by barcode (offs = first row of every barcode, row_bc = barcode of every row): one THREAD per
// row i pairs it with the later rows of its barcode, so a barcode with many contigs is spread over as many
// threads as it has rows.  Contigs are ordered by the host-supplied std::string rank (Arcs.cpp:1403); the
// orientation index is 2*(!Ahead)+(!Bhead) (:1418-1428).
__global__ void pair_kernel(const uint32_t* __restrict__ offs, const uint32_t* __restrict__ rows, const uint32_t* __restrict__ row_bc,
    uint32_t n_rows, const uint32_t* __restrict__ rank, unsigned long long* pmap, uint64_t pmap_mask, unsigned long long* pmap_count,
    uint64_t limit, uint32_t* overflow)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += gridDim.x * blockDim.x) {
		const uint32_t r1 = offs[row_bc[i] + 1];
		const uint32_t ri = rows[i];
		const uint32_t hi = ri >> 31;
		const uint32_t rki = rank[ri & 0x7FFFFFFFu];
		for (uint32_t j = i + 1; j < r1; ++j) {
			const uint32_t rj = rows[j];
			const uint32_t hj = rj >> 31;
			const uint32_t rkj = rank[rj & 0x7FFFFFFFu];
			if (rki == rkj)
				continue;
			const bool i_first = rki < rkj;
			const unsigned long long key = i_first ? ((unsigned long long)rki << 32) | rkj : ((unsigned long long)rkj << 32) | rki;
			const uint32_t ah = i_first ? hi : hj, bh = i_first ? hj : hi;
			if (!pmap_add(pmap, pmap_mask, pmap_count, limit, key, (ah ? 0u : 2u) + (bh ? 0u : 1u))) {
				*overflow = 1u;
				return;
			}
		}
	}
}

// occupied slots -> (key, slot index) records, arbitrary order
__global__ void pmap_collect_kernel(const unsigned long long* __restrict__ pmap, uint64_t cap, unsigned long long* __restrict__ keys,
    uint32_t* __restrict__ slots, unsigned long long* counter, uint64_t out_cap)
{
	for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < cap; s += (uint64_t)gridDim.x * blockDim.x) {
		const unsigned long long key = pmap[4 * s];
		if (key == kEmptyKey)
			continue;
		const unsigned long long i = atomicAdd(counter, 1ull);
		if (i < out_cap) {
			keys[i] = key;
			slots[i] = (uint32_t)s;
		}
	}
}

// counts of the sorted records, gathered from their slots
__global__ void pmap_gather_counts_kernel(const unsigned long long* __restrict__ pmap, const uint32_t* __restrict__ slots, uint64_t n,
    uint32_t* __restrict__ counts)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
	{
		// the four counters sit at byte 8 of the 32-byte slot: two 8-byte loads (a 16-byte one would be misaligned)
		const unsigned long long* p = pmap + 4 * (uint64_t)slots[i] + 1;
		const unsigned long long lo = p[0], hi = p[1];
		reinterpret_cast<uint4*>(counts)[i] = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
	}
}

// inv[rank[c]] = smallest contig index with that rank (contigs that share a name share a rank, and their
// hits are tallied under the first of them)
__global__ void rank_inverse_kernel(const uint32_t* __restrict__ rank, uint32_t n_contigs, uint32_t* __restrict__ inv)
{
	for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n_contigs; c += gridDim.x * blockDim.x)
		atomicMin(&inv[rank[c]], c);
}

// sorted keys -> contig indices of the two ends of every link
__global__ void pmap_names_kernel(const unsigned long long* __restrict__ keys, uint64_t n, const uint32_t* __restrict__ inv,
    uint32_t* __restrict__ a, uint32_t* __restrict__ b)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		const unsigned long long key = keys[i];
		a[i] = inv[(uint32_t)(key >> 32)];
		b[i] = inv[(uint32_t)key];
	}
}

// order-independent digest of the sorted map (for N-GPU invariance checks): digest[0] = sum of a 64-bit mix of
// every (key, counts) row, digest[1] = number of rows << 40 + sum of all counters
__global__ void pmap_digest_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ counts, uint64_t n,
    unsigned long long* digest)
{
	unsigned long long d0 = 0, d1 = 0;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint4 c = reinterpret_cast<const uint4*>(counts)[i];
		const unsigned long long lo = ((unsigned long long)c.y << 32) | c.x, hi = ((unsigned long long)c.w << 32) | c.z;
		d0 += mix64(keys[i] ^ mix64(lo + 0x9E3779B97F4A7C15ull) ^ (mix64(hi + 0xC2B2AE3D27D4EB4Full) << 1));
		d1 += (1ull << 40) + c.x + c.y + c.z + c.w;
	}
	for (int o = 16; o > 0; o >>= 1) {
		d0 += __shfl_down_sync(0xFFFFFFFFu, d0, o);
		d1 += __shfl_down_sync(0xFFFFFFFFu, d1, o);
	}
	if ((threadIdx.x & 31) == 0) {
		atomicAdd(digest, d0);
		atomicAdd(digest + 1, d1);
	}
}

__global__ void add_u32_kernel(uint32_t* __restrict__ acc, const uint32_t* __restrict__ x, uint64_t n)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		acc[i] += x[i];
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
