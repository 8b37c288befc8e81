// arks_map.cuh -- kernel 2: read-pair k-mer lookup + contig-end vote + barcode tally.
//
// Replaces the body of chromiumRead's parallel loop (Arcs/Arcs.cpp:1266-1292):
// checkReadSequence (:366-389) on both mates, bestContig (:939-1014) on both mates,
// and imap[barcode][contigRecord[c]]++ (:1280-1285).
//
// One warp per read pair.  The warp packs both mates to 2 bits in its private slice of
// shared memory (forward + reverse-complement streams + invalid-base mask); every lane
// then owns independent windows: it extracts the canonical key with funnel shifts
// (no rolling state), hashes it and probes the frozen table with one 16/32-byte
// non-allocating load.  Four probes per lane are in flight before the first is
// consumed.  Per-read votes live in registers, one tracked contig end per lane; the
// argmax (ties -> smallest contig end, as std::map iteration with a strict '<' gives,
// :996-1004) and the Jaccard gate (IEEE double division, :1006) finish in the warp.
#pragma once
#include "arks_device.cuh"

namespace arks {

constexpr int kMapWarps = 8; // warps per CTA
constexpr int kMapThreads = kMapWarps * 32;
constexpr int kRegionBases = 512;               // bases packed per region (32 lanes x 16)
constexpr int kRegionWords = kRegionBases / 16 + 6;
constexpr int kRegionInvWords = kRegionBases / 32 + 3;
constexpr int kProbeBatch = 4;                  // windows per lane in flight
constexpr int kMapMinBlocks = 2;                // CTAs per SM the register budget is sized for

struct MapCounters
{
	unsigned long long kmers_valid, kmers_invalid, found, recorded, dups, reads_pass, reads_fail, pairs_stored,
	    pairs_invalid, pairs_nogood, overflow;
};

struct WarpRegion
{
	uint32_t W[kRegionWords];
	uint32_t RC[kRegionWords];
	uint32_t INV[kRegionInvWords];
};

struct MapParams
{
	const uint8_t* table;
	uint64_t nslots;
	uint32_t k;
	uint64_t mask_hi, mask_lo;
	double j_index;
	const char* bases;
	const uint32_t* read_off;
	const uint32_t* barcode_id;
	uint32_t n_pairs;
	int32_t* conreci_out; // may be null
	const uint32_t* remap; // may be null
	uint32_t n_remap;
	// imap: open-address table of {u64 key = barcode<<32 | contig, u32 head, u32 tail}
	unsigned long long* imap;
	uint64_t imap_mask;
	unsigned long long* imap_count;
	MapCounters* ctr;
};

struct LaneStats
{
	uint32_t kv, ki, found, rec, dups;
};

// per-read vote state: lane i holds tracked contig end i
struct Track
{
	uint32_t c, cnt, n;
	bool overflow;
};

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
	x ^= x >> 33;
	x *= 0xFF51AFD7ED558CCDull;
	x ^= x >> 33;
	x *= 0xC4CEB9FE1A85EC53ull;
	x ^= x >> 33;
	return x;
}

// imap[barcode][contig end]++
__device__ __forceinline__ void
imap_add(unsigned long long* imap, uint64_t mask, unsigned long long* count, uint32_t barcode, uint32_t contig, uint32_t head_inc, uint32_t tail_inc)
{
	unsigned long long key = ((unsigned long long)barcode << 32) | contig;
	uint64_t slot = mix64(key) & mask;
	while (true) {
		unsigned long long* p = imap + 2 * slot;
		unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(p);
		if (cur == kEmptyKey) {
			cur = atomicCAS(p, (unsigned long long)kEmptyKey, key);
			if (cur == kEmptyKey) {
				atomicAdd(count, 1ull);
				cur = key;
			}
		}
		if (cur == key) {
			uint32_t* ht = reinterpret_cast<uint32_t*>(p + 1);
			if (head_inc)
				atomicAdd(ht, head_inc);
			if (tail_inc)
				atomicAdd(ht + 1, tail_inc);
			return;
		}
		slot = (slot + 1) & mask;
	}
}

// Packs a region of len <= kRegionBases bases.  Returns warp-uniform counts of N/n and of
// other invalid characters.
__device__ __forceinline__ void
warp_pack(WarpRegion& R, const char* src, uint32_t len, uint32_t lane, uint32_t& n_n, uint32_t& n_other)
{
	const uint32_t nwords = (len + 15) >> 4;
	uint32_t w = 0, inv16 = 0, nn = 0, no = 0;
	if (lane < nwords)
		w = pack_group(src, len, lane, &inv16, &nn, &no);
	R.W[lane] = w;
	uint32_t mirrored = __shfl_sync(0xFFFFFFFFu, w, (nwords - 1 - lane) & 31);
	if (lane < nwords)
		R.RC[lane] = rev2(~mirrored);
	uint32_t up = __shfl_down_sync(0xFFFFFFFFu, inv16, 1);
	if ((lane & 1) == 0)
		R.INV[lane >> 1] = inv16 | (up << 16);
	uint32_t any = __ballot_sync(0xFFFFFFFFu, inv16 != 0);
	n_n = any ? warp_sum(nn) : 0;
	n_other = any ? warp_sum(no) : 0;
	__syncwarp();
}

// checkReadSequence's character classes for a read of any length, without packing
__device__ __forceinline__ void
warp_classify(const char* src, uint32_t len, uint32_t lane, uint32_t& n_n, uint32_t& n_other)
{
	uint32_t nn = 0, no = 0;
	const uint32_t ngroups = (len + 15) >> 4;
	for (uint32_t g = lane; g < ngroups; g += 32) {
		uint32_t inv16, a, b;
		pack_group(src, len, g, &inv16, &a, &b);
		nn += a;
		no += b;
	}
	n_n = warp_sum(nn);
	n_other = warp_sum(no);
}

// checkReadSequence (Arcs.cpp:366-389): only ACGTN (any case), N fraction <= 0.02
__device__ __forceinline__ bool read_ok(uint32_t n_n, uint32_t n_other, uint32_t len)
{
	if (n_other)
		return false;
	double ar = (double)n_n / (double)len;
	return !(ar > 0.02);
}

__device__ __forceinline__ void track_add(Track& t, uint32_t lane, uint32_t c, uint32_t cnt)
{
	uint32_t has = __ballot_sync(0xFFFFFFFFu, lane < t.n && t.c == c);
	if (has) {
		if (lane == (uint32_t)__ffs(has) - 1)
			t.cnt += cnt;
	} else if (t.n < (uint32_t)kMaxTrack) {
		if (lane == t.n) {
			t.c = c;
			t.cnt = cnt;
		}
		t.n++;
	} else {
		t.overflow = true;
	}
}

// bestContig's window loop (Arcs.cpp:957-994) over windows [0, nw) of a packed region of
// Lc bases.
template <int KW>
__device__ __forceinline__ void
warp_windows(const WarpRegion& R, uint32_t Lc, uint32_t nw, const MapParams& P, uint32_t lane, Track& tr, LaneStats& st)
{
	const uint32_t Lp = ((Lc + 15) >> 4) << 4;
#pragma unroll 1
	for (uint32_t base = 0; base < nw; base += 32 * kProbeBatch) {
		Key128 key[kProbeBatch];
		uint64_t hi[kProbeBatch], lo[kProbeBatch];
		uint32_t val[kProbeBatch];
		uint32_t act = 0;
#pragma unroll
		for (int r = 0; r < kProbeBatch; ++r) {
			uint32_t p = base + r * 32 + lane;
			if (p < nw) {
				if (window_invalid(R.INV, p, P.k)) {
					st.ki++;
				} else {
					st.kv++;
					act |= 1u << r;
					key[r] = canonical_key<KW>(R.W, R.RC, p, P.k, Lp, P.mask_hi, P.mask_lo);
					load_slot<KW>(P.table, hash_to_slot(key_hash<KW>(key[r]), P.nslots), hi[r], lo[r], val[r]);
				}
			}
		}
		uint32_t hit[kProbeBatch];
#pragma unroll
		for (int r = 0; r < kProbeBatch; ++r) {
			hit[r] = 0;
			if (act & (1u << r)) {
				bool found = hi[r] == key[r].hi && lo[r] == key[r].lo;
				bool empty = hi[r] == kEmptyKey && (KW == 1 || lo[r] == kEmptyKey);
				if (!found && !empty) {
					// rare: the home slot holds another key -- walk the probe sequence
					uint64_t slot = hash_to_slot(key_hash<KW>(key[r]), P.nslots);
					do {
						slot = slot + 1 == P.nslots ? 0 : slot + 1;
						load_slot<KW>(P.table, slot, hi[r], lo[r], val[r]);
						found = hi[r] == key[r].hi && lo[r] == key[r].lo;
						empty = hi[r] == kEmptyKey && (KW == 1 || lo[r] == kEmptyKey);
					} while (!found && !empty);
				}
				if (found) {
					st.found++;
					if (val[r]) {
						st.rec++;
						hit[r] = val[r];
					} else {
						st.dups++;
					}
				}
			}
		}
		// merge this batch's hits into the per-read vote table
		while (true) {
			uint32_t m = 0xFFFFFFFFu;
#pragma unroll
			for (int r = 0; r < kProbeBatch; ++r)
				m = min(m, hit[r] ? hit[r] : 0xFFFFFFFFu);
			m = __reduce_min_sync(0xFFFFFFFFu, m);
			if (m == 0xFFFFFFFFu)
				break;
			uint32_t cnt = 0;
#pragma unroll
			for (int r = 0; r < kProbeBatch; ++r) {
				bool is = hit[r] == m;
				cnt += __popc(__ballot_sync(0xFFFFFFFFu, is));
				if (is)
					hit[r] = 0;
			}
			track_add(tr, lane, m, cnt);
		}
	}
}

// bestContig's window loop for a read longer than one region: repack chunk by chunk (cold path)
template <int KW>
__device__ __noinline__ void
warp_windows_long(WarpRegion& R, const char* src, uint32_t total, const MapParams& P, uint32_t lane, Track& tr, LaneStats& st)
{
	const uint32_t cw = kRegionBases - P.k + 1;
	for (uint32_t c0 = 0; c0 < total; c0 += cw) {
		uint32_t nwc = min(cw, total - c0);
		uint32_t Lc = nwc + P.k - 1;
		uint32_t a, b;
		__syncwarp();
		warp_pack(R, src + c0, Lc, lane, a, b);
		warp_windows<KW>(R, Lc, nwc, P, lane, tr, st);
	}
}

template <int KW>
__global__ void __launch_bounds__(kMapThreads, kMapMinBlocks) map_pairs_kernel(MapParams P)
{
	__shared__ WarpRegion regions[kMapWarps][2];
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warp = threadIdx.x >> 5;
	WarpRegion* R = regions[warp];
	LaneStats st{0, 0, 0, 0, 0};
	uint32_t pass = 0, fail = 0, stored = 0, invalid = 0, nogood = 0;
	bool overflow = false;
	const uint32_t nwarps = gridDim.x * kMapWarps;
#pragma unroll 1
	for (uint32_t pair = blockIdx.x * kMapWarps + warp; pair < P.n_pairs; pair += nwarps) {
		const uint32_t o0 = P.read_off[2 * pair], o1 = P.read_off[2 * pair + 1], o2 = P.read_off[2 * pair + 2];
		const uint32_t l1 = o1 - o0, l2 = o2 - o1;
		const bool shortpair = l1 <= (uint32_t)kRegionBases && l2 <= (uint32_t)kRegionBases;
		uint32_t nn1, no1, nn2, no2;
		__syncwarp();
		if (shortpair) {
			warp_pack(R[0], P.bases + o0, l1, lane, nn1, no1);
			warp_pack(R[1], P.bases + o1, l2, lane, nn2, no2);
		} else {
			warp_classify(P.bases + o0, l1, lane, nn1, no1);
			warp_classify(P.bases + o1, l2, lane, nn2, no2);
		}
		uint32_t c[2] = {0, 0};
		if (read_ok(nn1, no1, l1) && read_ok(nn2, no2, l2)) {
#pragma unroll 1
			for (int rd = 0; rd < 2; ++rd) {
				// bestContig (Arcs.cpp:939-1014) for mate rd
				const uint32_t len = rd ? l2 : l1;
				const uint32_t total = len >= P.k ? len - P.k + 1 : 0;
				Track tr{0, 0, 0, false};
				if (total) {
					if (shortpair)
						warp_windows<KW>(R[rd], len, total, P, lane, tr, st);
					else
						warp_windows_long<KW>(R[0], P.bases + (rd ? o1 : o0), total, P, lane, tr, st);
				}
				// argmax count, ties -> smallest contig end
				uint32_t mycnt = lane < tr.n ? tr.cnt : 0;
				uint32_t best_cnt = __reduce_max_sync(0xFFFFFFFFu, mycnt);
				uint32_t cand = (best_cnt && lane < tr.n && tr.cnt == best_cnt) ? tr.c : 0xFFFFFFFFu;
				uint32_t best_c = __reduce_min_sync(0xFFFFFFFFu, cand);
				double maxj = best_cnt ? (double)best_cnt / (double)total : 0.0;
				overflow |= tr.overflow;
				if (maxj > P.j_index) {
					pass++;
					c[rd] = best_cnt ? best_c : 0;
				} else {
					fail++;
				}
			}
		} else {
			invalid++;
		}
		uint32_t out = 0;
		if (c[0] != 0 && c[0] == c[1]) {
			stored++;
			out = c[0];
			if (lane == 0) {
				uint32_t cc = (P.remap && out < P.n_remap) ? P.remap[out] : out;
				imap_add(P.imap, P.imap_mask, P.imap_count, P.barcode_id[pair], (cc - 1) >> 1, (cc & 1u), (cc & 1u) ^ 1u);
			}
		} else {
			nogood++;
		}
		if (P.conreci_out && lane == 0)
			P.conreci_out[pair] = (int32_t)out;
	}
	// flush counters: lane-private k-mer counters are summed over the warp, the per-pair /
	// per-read counters are warp-uniform
	uint32_t kv = warp_sum(st.kv), ki = warp_sum(st.ki), fo = warp_sum(st.found), re = warp_sum(st.rec),
	         du = warp_sum(st.dups);
	if (lane == 0) {
		MapCounters* ctr = P.ctr;
		if (kv) atomicAdd(&ctr->kmers_valid, (unsigned long long)kv);
		if (ki) atomicAdd(&ctr->kmers_invalid, (unsigned long long)ki);
		if (fo) atomicAdd(&ctr->found, (unsigned long long)fo);
		if (re) atomicAdd(&ctr->recorded, (unsigned long long)re);
		if (du) atomicAdd(&ctr->dups, (unsigned long long)du);
		if (pass) atomicAdd(&ctr->reads_pass, (unsigned long long)pass);
		if (fail) atomicAdd(&ctr->reads_fail, (unsigned long long)fail);
		if (stored) atomicAdd(&ctr->pairs_stored, (unsigned long long)stored);
		if (invalid) atomicAdd(&ctr->pairs_invalid, (unsigned long long)invalid);
		if (nogood) atomicAdd(&ctr->pairs_nogood, (unsigned long long)nogood);
		if (overflow) atomicAdd(&ctr->overflow, 1ull);
	}
}

} // namespace arks
