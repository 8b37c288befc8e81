// arks_index.cuh -- kernel 1: contig-end k-merisation into the exact GPU hash table.
//
// Replaces mapKmers (Arcs/Arcs.cpp:869-929) + ReadsProcessor::prepSeq + the
// google::sparse_hash_map / CityHash container (Arcs/Arcs.h:140-158).
//
// Per batch of contig ends (ASCII, concatenated):
//   inv_mask_kernel   1 bit per base: not ACGTacgt                     (streaming)
//   walk_kernel       one thread per end resolves mapKmers' sequential walk: a NULL
//                     window at i jumps to i+k (:922-925), which can skip windows that
//                     would have been valid; those are flagged in a skip bitmask and
//                     the NULL events are counted.
//   insert_kernel     one CTA per tile of 1024 windows: pack the tile's bases to 2 bits
//                     in shared memory, every thread extracts canonical keys for its
//                     windows and inserts them (128-bit CAS claim of the slot, then a
//                     64-bit CAS on the slot's bookkeeping word).
//   finalize_kernel   collapses bookkeeping into the final value per key + counters.
//
// Bookkeeping word w per key while building: (min conreci << 32) | multi-flag << 31 |
// occurrences at that min conreci.  The reference's value rule (first conreci; any other
// conreci -> 0, sticky; :903-920) is order independent: value = multi ? 0 : min conreci.
// s_numkmersremdup counts, for ends processed in increasing conreci order (as
// getContigKmers does), every occurrence of a key that is not in its smallest conreci,
// so removed = sum over keys (total - count_at_min) = kmers_valid - sum count_at_min.
#pragma once
#include "arks_device.cuh"

namespace arks {

constexpr int kTileWindows = 1024;
constexpr int kInsertThreads = 256;
constexpr int kInsertBatch = kTileWindows / kInsertThreads; // windows per thread and tile
#ifndef ARKS_INSERT_MIN_BLOCKS
#define ARKS_INSERT_MIN_BLOCKS 3
#endif
constexpr int kInsertMinBlocks = ARKS_INSERT_MIN_BLOCKS; // CTAs per SM the register budget is sized for
// words of packed region per tile: (1024 + 64 - 1 + 15)/16 = 68, +5 slack for extraction
constexpr int kTileWords = 68 + 6;
constexpr int kTileInvWords = 35 + 3;

struct IndexCounters
{
	unsigned long long kmers_valid, kmers_null, recorded, unique, sum_cmin, probe_fail;
};

// ---- 1 bit per base: invalid --------------------------------------------------------
__global__ void inv_mask_kernel(const char* __restrict__ bases, uint64_t n_bases, uint32_t* __restrict__ inv)
{
	uint64_t n_words = (n_bases + 31) / 32;
	for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t b0 = w * 32;
		uint32_t remain = (uint32_t)min((uint64_t)32, n_bases - b0);
		uint32_t m = 0;
#pragma unroll
		for (int g = 0; g < 2; ++g) {
			if (remain > 16u * g) {
				uint32_t inv16, nn, no;
				pack_group(bases + b0 + 16 * g, remain - 16 * g, 0, &inv16, &nn, &no);
				m |= inv16 << (16 * g);
			}
		}
		inv[w] = m;
	}
}

// first set bit of the LSB-first bitmask in [from, to), or `to`
__device__ __forceinline__ uint64_t next_set_bit(const uint32_t* mask, uint64_t from, uint64_t to)
{
	if (from >= to)
		return to;
	uint64_t w = from >> 5;
	uint32_t cur = mask[w] & (0xFFFFFFFFu << (from & 31));
	uint64_t wend = (to - 1) >> 5;
	while (true) {
		if (cur) {
			uint64_t pos = (w << 5) + (uint64_t)(__ffs(cur) - 1);
			return pos < to ? pos : to;
		}
		if (w == wend)
			return to;
		cur = mask[++w];
	}
}

// ---- mapKmers' walk (Arcs.cpp:886-927) -------------------------------------------------
__global__ void walk_kernel(const uint64_t* __restrict__ end_off, uint32_t n_ends, uint32_t k,
    const uint32_t* __restrict__ inv, uint32_t* __restrict__ skip, IndexCounters* ctr)
{
	uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long nnull = 0;
	if (e < n_ends) {
		uint64_t s = end_off[e], t = end_off[e + 1];
		uint64_t len = t - s;
		if (len >= k) {
			uint64_t last = len - k; // last window start
			uint64_t i = 0;
			while (i <= last) {
				uint64_t q = next_set_bit(inv, s + i, t) - s;
				if (q >= len)
					break; // every remaining window is valid
				uint64_t ip = (q + 1 >= k && q + 1 - k > i) ? q + 1 - k : i; // first NULL window reached
				if (ip > last)
					break;
				nnull++;
				// the jump skips windows [ip, ip+k); flag them (those containing q are NULL anyway)
				uint64_t a = s + ip, b = min(s + ip + k, s + last + 1);
				for (uint64_t w = a >> 5; w <= (b - 1) >> 5; ++w) {
					uint64_t lo = max(a, w << 5), hi = min(b, (w + 1) << 5);
					uint32_t m = (hi - lo) >= 32 ? 0xFFFFFFFFu : (((1u << (hi - lo)) - 1u) << (lo & 31));
					atomicOr(&skip[w], m);
				}
				i = ip + k;
			}
		}
	}
	nnull = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)nnull); // per-thread counts are small
	if ((threadIdx.x & 31) == 0 && nnull)
		atomicAdd(&ctr->kmers_null, nnull);
}

// ---- insert --------------------------------------------------------------------------
__device__ __forceinline__ void
cas128(uint64_t* p, uint64_t c_lo, uint64_t c_hi, uint64_t n_lo, uint64_t n_hi, uint64_t& o_lo, uint64_t& o_hi)
{
	asm volatile("{\n\t.reg .b128 c, n, d;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 n, {%4, %5};\n\t"
	             "atom.global.cas.b128 d, [%6], c, n;\n\tmov.b128 {%0, %1}, d;\n\t}"
	             : "=l"(o_lo), "=l"(o_hi)
	             : "l"(c_lo), "l"(c_hi), "l"(n_lo), "l"(n_hi), "l"(p)
	             : "memory");
}

// the bookkeeping word after one more occurrence of its key in contig end `conreci`
__device__ __forceinline__ unsigned long long note_next(unsigned long long old, uint32_t conreci)
{
	if (old == kEmptyW)
		return ((unsigned long long)conreci << 32) | 1ull;
	const uint32_t mc = (uint32_t)(old >> 32);
	if (conreci < mc)
		return ((unsigned long long)conreci << 32) | kMultiFlag | 1ull;
	if (conreci == mc)
		return old + 1ull;
	return old | kMultiFlag;
}

// `old` was read from memory (by a failed compare-and-swap): retry until the occurrence is in
__device__ __forceinline__ void note_occurrence_from(unsigned long long* wp, uint32_t conreci, unsigned long long old)
{
	while (true) {
		const unsigned long long nw = note_next(old, conreci);
		if (nw == old)
			return;
		const unsigned long long prev = atomicCAS(wp, old, nw);
		if (prev == old)
			return;
		old = prev;
	}
}

// continues the probe sequence behind `slot` (whose first compare-and-swap met another key)
template <int KW>
__device__ __noinline__ uint64_t* find_or_claim_behind(uint8_t* table, uint64_t nslots, uint64_t slot, uint64_t key_hi, uint64_t key_lo, bool* claimed)
{
	const Key128 key{key_hi, key_lo};
	*claimed = false;
	for (uint64_t probes = 1; probes < nslots; ++probes) {
		slot = slot + 1 == nslots ? 0 : slot + 1;
		uint64_t* p = reinterpret_cast<uint64_t*>(table + slot * kSlotBytes);
		uint64_t hi, lo;
		if (KW == 1) {
			hi = atomicCAS(reinterpret_cast<unsigned long long*>(p), (unsigned long long)kEmptyKey, (unsigned long long)key.hi);
			lo = hi == kEmptyKey ? kEmptyKey : key.lo;
		} else {
			cas128(p, kEmptyKey, kEmptyKey, key.hi, key.lo, hi, lo);
		}
		if (hi == kEmptyKey && lo == kEmptyKey) {
			*claimed = true;
			return p;
		}
		if (hi == key.hi && lo == key.lo)
			return p;
	}
	return nullptr;
}

struct IndexTile
{
	uint32_t end;    // index of the contig end in this batch
	uint32_t start;  // first window of the tile within the end (multiple of kTileWindows)
};

// Persistent packed copy of every contig end (global base coordinates; every end starts at a
// multiple of 32): T = 2-bit text, 16 bases per word, MSB first; TINS bit g = the window
// starting at g was inserted by mapKmers' walk; TUNIQ bit g = its key maps to one contig end.
struct ContigText
{
	uint32_t* T;
	uint32_t* TINS;
	uint32_t* TUNIQ;
	const uint64_t* end_g0;   // per global end index
	const uint32_t* end_len;
	const uint32_t* end_cr;
	uint64_t n_bases;         // coordinates in use
};

template <int KW>
__global__ void __launch_bounds__(kInsertThreads, kInsertMinBlocks)
insert_kernel(const IndexTile* __restrict__ tiles, uint32_t n_tiles, const char* __restrict__ bases,
    const uint64_t* __restrict__ end_off, const uint32_t* __restrict__ conreci, const uint32_t* __restrict__ skip,
    const uint64_t* __restrict__ batch_g0, uint32_t first_global_end, ContigText ct, uint8_t* table, uint64_t nslots, uint32_t k,
    uint64_t mask_hi, uint64_t mask_lo, IndexCounters* ctr)
{
	__shared__ uint32_t W[kTileWords];
	__shared__ uint32_t RC[kTileWords];
	__shared__ uint32_t INV[kTileInvWords];
	__shared__ uint32_t inv16s[kTileWords];
	unsigned long long my_valid = 0;
	bool fail = false;
	for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
		IndexTile tl = tiles[tile];
		uint64_t s = end_off[tl.end];
		uint32_t len = (uint32_t)(end_off[tl.end + 1] - s);
		uint32_t nwin_total = len - k + 1;
		uint32_t nw = min((uint32_t)kTileWindows, nwin_total - tl.start);
		uint32_t Lc = nw + k - 1;
		uint32_t nwords = (Lc + 15) >> 4;
		const bool last_tile = tl.start + nw == nwin_total;
		const char* src = bases + s + tl.start;
		const uint64_t g0 = batch_g0[tl.end] + tl.start; // global coordinate of the tile's first base (multiple of 32)
		for (uint32_t g = threadIdx.x; g < nwords; g += blockDim.x) {
			uint32_t inv16, nn, no;
			uint32_t w = pack_group(src, Lc, g, &inv16, &nn, &no);
			W[g] = w;
			inv16s[g] = inv16;
			// persistent text: a tile owns its first kTileWindows bases; the end's last tile owns the rest
			if (last_tile || g < (uint32_t)kTileWindows / 16)
				ct.T[(g0 >> 4) + g] = w;
		}
		__syncthreads();
		for (uint32_t g = threadIdx.x; g < nwords; g += blockDim.x)
			RC[g] = rev2(~W[nwords - 1 - g]);
		for (uint32_t m = threadIdx.x; m < (nwords + 1) / 2; m += blockDim.x)
			INV[m] = inv16s[2 * m] | ((2 * m + 1 < nwords ? inv16s[2 * m + 1] : 0u) << 16);
		__syncthreads();
		const uint32_t cr = conreci[tl.end];
		const uint64_t gpos0 = s + tl.start;
		const uint64_t end_tag = (uint64_t)(first_global_end + tl.end) << 40;
		// Every thread takes kInsertBatch windows of the tile and walks them through the insert in STAGES, so that its
		// memory round trips overlap instead of forming one dependent chain per window (the kernel was latency-bound:
		// issue-active 14 %, DRAM 16 %): all keys -> all first compare-and-swaps (the probe IS the claim: it returns
		// what the slot holds) -> all bookkeeping compare-and-swaps (guessing the word: unset after a claim, "seen once
		// in this contig end" otherwise) -> the few that need another round.
		Key128 key[kInsertBatch];
		uint64_t slot[kInsertBatch], ohi[kInsertBatch], olo[kInsertBatch];
		uint32_t act = 0, fcm = 0;
#pragma unroll
		for (int r = 0; r < kInsertBatch; ++r) {
			const uint32_t p = r * kInsertThreads + threadIdx.x;
			key[r] = Key128{0, 0};
			slot[r] = 0;
			if (p < nw) {
				const uint64_t gp = gpos0 + p;
				const bool skipped = (skip[gp >> 5] >> (gp & 31)) & 1u;
				if (!skipped && !window_invalid(INV, p, k)) {
					bool fwd_canon;
					key[r] = canonical_key<KW>(W, RC, p, k, nwords * 16, mask_hi, mask_lo, &fwd_canon);
					slot[r] = hash_to_slot(key_hash<KW>(key[r]), nslots);
					act |= 1u << r;
					fcm |= (uint32_t)fwd_canon << r;
				}
			}
		}
#pragma unroll
		for (int r = 0; r < kInsertBatch; ++r) {
			ohi[r] = olo[r] = 0;
			if (act & (1u << r)) {
				uint64_t* p = reinterpret_cast<uint64_t*>(table + slot[r] * kSlotBytes);
				if (KW == 1) {
					ohi[r] = atomicCAS(reinterpret_cast<unsigned long long*>(p), (unsigned long long)kEmptyKey, (unsigned long long)key[r].hi);
					olo[r] = ohi[r] == kEmptyKey ? kEmptyKey : 0ull;
				} else {
					// memory order of the pair is (p[0], p[1]) = (hi, lo); b128 = {low 64, high 64}
					cas128(p, kEmptyKey, kEmptyKey, key[r].hi, key[r].lo, ohi[r], olo[r]);
				}
			}
		}
		uint64_t* sp[kInsertBatch];
		unsigned long long guess[kInsertBatch], prev[kInsertBatch];
#pragma unroll
		for (int r = 0; r < kInsertBatch; ++r) {
			sp[r] = nullptr;
			guess[r] = prev[r] = 0;
			if (act & (1u << r)) {
				bool claimed = ohi[r] == kEmptyKey && olo[r] == kEmptyKey;
				const bool found = ohi[r] == key[r].hi && (KW == 1 || olo[r] == key[r].lo);
				sp[r] = reinterpret_cast<uint64_t*>(table + slot[r] * kSlotBytes);
				if (!claimed && !found)
					sp[r] = find_or_claim_behind<KW>(table, nslots, slot[r], key[r].hi, key[r].lo, &claimed);
				if (sp[r] == nullptr) {
					fail = true;
					act &= ~(1u << r);
				} else {
					if (claimed)
						sp[r][3] = end_tag | ((uint64_t)((fcm >> r) & 1u) << kPosBits) | (g0 + r * kInsertThreads + threadIdx.x);
					guess[r] = claimed ? (unsigned long long)kEmptyW : (((unsigned long long)cr << 32) | 1ull);
					prev[r] = atomicCAS(reinterpret_cast<unsigned long long*>(sp[r] + 2), guess[r], note_next(guess[r], cr));
					my_valid++;
				}
			}
		}
#pragma unroll
		for (int r = 0; r < kInsertBatch; ++r) {
			if ((act & (1u << r)) && prev[r] != guess[r])
				note_occurrence_from(reinterpret_cast<unsigned long long*>(sp[r] + 2), cr, prev[r]);
			// 32 consecutive windows of one warp = one word of the inserted-window mask
			const uint32_t word = __ballot_sync(0xFFFFFFFFu, (act >> r) & 1u);
			const uint32_t pb = r * kInsertThreads;
			if ((threadIdx.x & 31) == 0 && pb + (threadIdx.x & ~31u) < nw)
				ct.TINS[(g0 + pb + threadIdx.x) >> 5] = word;
		}
		__syncthreads();
	}
	uint32_t v = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)my_valid);
	if ((threadIdx.x & 31) == 0 && v)
		atomicAdd(&ctr->kmers_valid, (unsigned long long)v);
	if (fail)
		atomicAdd(&ctr->probe_fail, 1ull);
}

// ---- finalize ------------------------------------------------------------------------
// One pass over the table: collapses the build bookkeeping of every key into its final value (+ the counters),
// adds the key to the membership filter (bloom_probe in arks_device.cuh; keys whose value is 0 are included: a
// lookup that finds them counts as "found", Arcs.cpp:969-971), and marks the text window that claimed the slot
// as "unique" when the key maps to one contig end (the other occurrences of such a key -- rare -- are settled by
// uniq_mask_kernel).  Slots already finalised (high half of the word 0) are left alone.
template <int KW>
__global__ void finalize_kernel(uint8_t* table, uint64_t nslots, IndexCounters* ctr, unsigned long long* bloom, uint64_t bloom_words,
    uint32_t* TUNIQ)
{
	unsigned long long rec = 0, uniq = 0, cmin = 0;
	for (uint64_t sidx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; sidx < nslots; sidx += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t* p = reinterpret_cast<uint64_t*>(table + sidx * kSlotBytes);
		const ulonglong4 sl = *reinterpret_cast<const ulonglong4*>(p);
		if (slot_empty<KW>(sl.x, sl.y))
			continue;
		const uint64_t w = sl.z;
		if ((w >> 32) == 0)
			continue; // already final
		uint32_t mc = (uint32_t)(w >> 32);
		bool multi = (w & kMultiFlag) != 0;
		p[2] = multi ? 0ull : (uint64_t)mc;
		rec++;
		uniq += multi ? 0 : 1;
		cmin += (uint32_t)w & 0x7FFFFFFFu;
		if (bloom) {
			const Key128 key{sl.x, KW == 2 ? sl.y : 0ull};
			const BloomProbe b = bloom_probe(key_hash<KW>(key), bloom_words);
			atomicOr(bloom + b.word, ((unsigned long long)bloom_mask_hi(b.sel) << 32) | bloom_mask_lo(b.sel));
		}
		if (!multi && TUNIQ) {
			const uint64_t g = sl.w & kPosMask;
			atomicOr(TUNIQ + (g >> 5), 1u << (g & 31u));
		}
	}
	// block reduce through warp sums
	for (int o = 16; o > 0; o >>= 1) {
		rec += __shfl_down_sync(0xFFFFFFFFu, rec, o);
		uniq += __shfl_down_sync(0xFFFFFFFFu, uniq, o);
		cmin += __shfl_down_sync(0xFFFFFFFFu, cmin, o);
	}
	if ((threadIdx.x & 31) == 0) {
		if (rec)
			atomicAdd(&ctr->recorded, rec);
		if (uniq)
			atomicAdd(&ctr->unique, uniq);
		if (cmin)
			atomicAdd(&ctr->sum_cmin, cmin);
	}
}

// frozen-table lookup of one key: value, or kMiss
template <int KW>
__device__ __forceinline__ uint32_t lookup_value(const uint8_t* table, uint64_t nslots, const Key128& key)
{
	uint64_t slot = hash_to_slot(key_hash<KW>(key), nslots);
	while (true) {
		uint64_t hi, lo, posinfo;
		uint32_t val;
		load_slot(table, slot, hi, lo, val, posinfo);
		if (slot_matches<KW>(hi, lo, key))
			return val;
		if (slot_empty<KW>(hi, lo))
			return kMiss;
		slot = slot + 1 == nslots ? 0 : slot + 1;
	}
}

// TUNIQ bit g = the key of the inserted window at g maps to exactly one contig end
template <int KW>
__global__ void uniq_mask_kernel(ContigText ct, const uint8_t* table, uint64_t nslots, uint32_t k, uint64_t mask_hi, uint64_t mask_lo)
{
	const uint64_t n_words = (ct.n_bases + 31) >> 5;
	const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
	const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	const uint32_t lane = threadIdx.x & 31;
	for (uint64_t w = warp0; w < n_words; w += nwarps) {
		// windows already marked by finalize_kernel (the occurrence that claimed its key's slot) need no lookup
		const uint32_t have = ct.TUNIQ[w];
		const uint32_t ins = ct.TINS[w] & ~have;
		if (ins == 0u)
			continue;
		bool uq = false;
		if ((ins >> lane) & 1u) {
			const uint64_t g = (w << 5) + lane;
			Key128 f = extract_window<KW>(ct.T + (g >> 4), (uint32_t)(g & 15), mask_hi, mask_lo);
			bool fc;
			Key128 key = canonical_from_forward<KW>(f, k, &fc);
			uint32_t v = lookup_value<KW>(table, nslots, key);
			uq = v != 0 && v != kMiss;
		}
		const uint32_t word = __ballot_sync(0xFFFFFFFFu, uq);
		if (lane == 0 && word)
			ct.TUNIQ[w] = have | word;
	}
}

// copies (key, value) of every occupied slot to dense arrays (for tests / dumps)
template <int KW>
__global__ void dump_kernel(const uint8_t* table, uint64_t nslots, uint64_t* keys_hi, uint64_t* keys_lo, int32_t* vals,
    unsigned long long* counter, uint64_t cap)
{
	for (uint64_t sidx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; sidx < nslots; sidx += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t* p = reinterpret_cast<const uint64_t*>(table + sidx * kSlotBytes);
		if (slot_empty<KW>(p[0], p[1]))
			continue;
		unsigned long long i = atomicAdd(counter, 1ull);
		if (i < cap) {
			keys_hi[i] = p[0];
			keys_lo[i] = KW == 2 ? p[1] : 0ull;
			vals[i] = (int32_t)(uint32_t)p[2];
		}
	}
}

} // namespace arks
