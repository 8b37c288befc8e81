// arks_map.cuh -- kernel 2: read-pair k-mer lookup + contig-end vote + barcode tally.
//
// Replaces the body of chromiumRead's parallel loop (Arcs/Arcs.cpp:1266-1292):
// checkReadSequence (:366-389) on both mates, bestContig (:939-1014) on both mates,
// and imap[barcode][contigRecord[c]]++ (:1280-1285).
//
// Two kernels per batch:
//  * map_groups_kernel -- a warp takes 16 pairs = 32 reads: packs them to 2 bits, then ONE LANE PER READ
//    probes a few seed windows, compares the read with the packed contig text along the seed's
//    diagonal and classifies every window word-parallel (found iff the text window was inserted,
//    recorded iff its key is unique); the windows that still have to be looked up (they overlap a
//    sequencing error) are dealt flat over the warp in chunks with a rolling key, go through an
//    L2-resident membership filter first, and reach the table only on a filter positive.
//  * map_slow_kernel -- the general warp-per-pair path (per-window test, compacted probe list, votes for
//    up to 32 contig ends in registers) for the pairs the first kernel hands over: no seed hit, votes
//    for three contig ends, reads longer than 256 bases.
// map_pairs_kernel (warp per pair for every pair) is kept as an independent implementation for A/B
// runs and tests/test_gpu_properties.py.  The argmax (ties -> smallest contig end, as std::map
// iteration with a strict '<' gives, :996-1004) and the Jaccard gate (:1006) are exact: the gate is a
// host-built integer table of the reference's double-precision expression.
#pragma once
#include "arks_device.cuh"
#include <cstdio>

namespace arks {

constexpr int kMapWarps = 8; // warps per CTA
constexpr int kMapThreads = kMapWarps * 32;
constexpr int kRegionBases = 512;               // bases packed per region (32 lanes x 16)
constexpr int kRegionWords = kRegionBases / 16 + 6;
constexpr int kRegionInvWords = kRegionBases / 32 + 3;
constexpr int kMaskWords = kRegionBases / 32 + 2;          // words of TINS / TUNIQ a region can span
constexpr int kMWords = kRegionInvWords + 2 * kMaskWords;  // per-warp scratch: mismatch mask + the two window masks
#ifndef ARKS_PROBE_BATCH
#define ARKS_PROBE_BATCH 1
#endif
#ifndef ARKS_MAP_MIN_BLOCKS
#define ARKS_MAP_MIN_BLOCKS 5
#endif
#ifndef ARKS_SEEDS
#define ARKS_SEEDS 2
#endif
constexpr int kProbeBatch = ARKS_PROBE_BATCH;                  // windows per lane in flight
constexpr int kSeeds = ARKS_SEEDS;                       // seed windows probed per mate
constexpr int kMapMinBlocks = ARKS_MAP_MIN_BLOCKS;                // CTAs per SM the register budget is sized for

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
	// packed contig text for seed-and-extend (see ContigText in arks_index.cuh)
	const uint32_t* ct_T;
	const uint32_t* ct_TINS;
	const uint32_t* ct_TUNIQ;
	const uint64_t* ct_end_g0;
	const uint32_t* ct_end_len;
	const uint32_t* ct_end_cr;
	uint64_t ct_n_bases;
	int use_extension;
	// exact integer forms of the two double-precision tests, built on the host with the
	// reference's expressions (both tables have kRegionBases + 1 entries):
	//   jmin[total] = smallest count with (double)count / (double)total > j_index (UINT32_MAX: none)
	//   nmax[len]   = largest number of Ns with !((double)n / (double)len > 0.02)
	const uint32_t* jmin;
	const uint32_t* nmax;
	// pairs the group kernel could not finish: one 64-byte WorkRecord each
	struct WorkRecord* work;
	uint32_t* work_count;
	// membership prefilter over the table's keys (bloom_probe, arks_device.cuh); may be null
	const unsigned long long* bloom;
	uint64_t bloom_words;
	int lane_general; // 0: the group kernel finishes only reads that equal the contig text (A/B runs)
};

// Everything map_slow_kernel needs to start on a deferred pair, in one aligned 64-byte record (so the
// record of the pair after next can be prefetched by address alone, and the next pair's record is
// an L1 hit from which ITS bases / contig text / masks are prefetched while the current pair runs).
struct __align__(16) WorkRecord
{
	uint32_t pair;
	uint32_t off[3];    // read_off[2*pair .. 2*pair+2]
	uint32_t state[2];  // per mate: kMateSlow / kMateUnknown / the contig end already decided
	uint32_t meta[2];   // per mate: bit 0 seed hit, bit 1 seed's canonical key is the read's forward strand, bits 8.. seed window
	uint32_t seed_lo[2], seed_hi[2]; // posinfo of the seed's slot
	uint32_t pad[4];
};

// mate_state values: a contig end (or 0) decided by the group kernel, or one of
constexpr uint32_t kMateSlow = 0xFFFFFFFFu;    // valid pair, this mate still has to be resolved
constexpr uint32_t kMateUnknown = 0xFFFFFFFEu; // nothing has been done for this pair yet

struct LaneStats
{
	uint32_t kv, ki, found, rec, dups;
};

// per-read vote state: lane i holds tracked contig end i.  A read that votes for more than kMaxTrack contig
// ends (uncut long reads, reads over many short contigs) is counted in several passes over its windows: pass
// `part` of `parts` (a power of two) only sees the contig ends whose hash falls into it.
struct Track
{
	uint32_t c, cnt, n;
	bool overflow;
	uint32_t part, parts;
};

__device__ __forceinline__ bool track_takes(const Track& t, uint32_t c)
{
	return t.parts <= 1u || (((c * 0x9E3779B1u) >> 16) & (t.parts - 1u)) == t.part;
}

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

__device__ __noinline__ void
imap_add_call(unsigned long long* imap, uint64_t mask, unsigned long long* count, uint32_t barcode, uint32_t contig, uint32_t head_inc, uint32_t tail_inc)
{
	imap_add(imap, mask, count, barcode, contig, head_inc, tail_inc);
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

// Packs both mates in one pass (each <= 256 bases): lanes 0-15 pack mate 0, lanes 16-31
// mate 1.  Returns, on every lane, the N / other-invalid counts of both mates.
__device__ __forceinline__ void
warp_pack2(WarpRegion* R, const char* s0, uint32_t l0, const char* s1, uint32_t l1, uint32_t lane, uint32_t& nn0, uint32_t& no0,
    uint32_t& nn1, uint32_t& no1)
{
	const uint32_t half = lane >> 4, j = lane & 15u;
	const uint32_t len = half ? l1 : l0;
	const char* src = half ? s1 : s0;
	const uint32_t nwords = (len + 15) >> 4;
	uint32_t w = 0, inv16 = 0, nn = 0, no = 0;
	if (j < nwords)
		w = pack_group(src, len, j, &inv16, &nn, &no);
	WarpRegion& Rr = R[half];
	Rr.W[j] = w;
	const uint32_t mirrored = __shfl_sync(0xFFFFFFFFu, w, (half << 4) | ((nwords - 1 - j) & 15u));
	if (j < nwords)
		Rr.RC[j] = rev2(~mirrored);
	const uint32_t up = __shfl_down_sync(0xFFFFFFFFu, inv16, 1);
	if ((j & 1) == 0)
		Rr.INV[j >> 1] = inv16 | (up << 16);
	const uint32_t any = __ballot_sync(0xFFFFFFFFu, inv16 != 0);
	nn0 = no0 = nn1 = no1 = 0;
	if (any) {
		// pack (n, other) of each half into one word per lane and add across the warp
		const uint32_t v = nn | (no << 16);
		const uint32_t lo = __reduce_add_sync(0xFFFFFFFFu, half ? 0u : v);
		const uint32_t hi = __reduce_add_sync(0xFFFFFFFFu, half ? v : 0u);
		nn0 = lo & 0xFFFFu;
		no0 = lo >> 16;
		nn1 = hi & 0xFFFFu;
		no1 = hi >> 16;
	}
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
__device__ __forceinline__ bool read_ok(uint32_t n_n, uint32_t n_other, uint32_t len, const uint32_t* nmax)
{
	if (n_other)
		return false;
	if (n_n == 0)
		return true; // 0/len = 0 (or NaN for len 0): never > 0.02
	if (len <= (uint32_t)kRegionBases)
		return n_n <= nmax[len];
	double ar = (double)n_n / (double)len;
	return !(ar > 0.02);
}

__device__ __forceinline__ void track_add(Track& t, uint32_t lane, uint32_t c, uint32_t cnt)
{
	if (!track_takes(t, c)) // warp-uniform: c is
		return;
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

// ---------------------------------------------------------------------------------------
// Probing a list of windows of one packed region (bestContig's loop body, Arcs.cpp:957-994,
// for the windows that could not be resolved by extension).  `list` == nullptr means the
// windows are 0..n-1 themselves.  Does NOT count kv/ki (the caller classified the windows).
// ---------------------------------------------------------------------------------------
#ifdef ARKS_PROBE_NOINLINE
#define ARKS_PROBE_ATTR __noinline__
#else
#define ARKS_PROBE_ATTR __forceinline__
#endif
template <int KW>
__device__ ARKS_PROBE_ATTR void
warp_probe_windows(const WarpRegion& R, uint32_t Lc, const uint16_t* list, uint32_t n, const MapParams& P, uint32_t lane,
    Track& tr, LaneStats& st)
{
	const uint32_t Lp = ((Lc + 15) >> 4) << 4;
#pragma unroll 1
	for (uint32_t base = 0; base < n; base += 32 * kProbeBatch) {
		Key128 key[kProbeBatch];
		uint64_t hi[kProbeBatch], lo[kProbeBatch], pi;
		uint32_t val[kProbeBatch];
		uint32_t act = 0;
#pragma unroll
		for (int r = 0; r < kProbeBatch; ++r) {
			uint32_t i = base + r * 32 + lane;
			if (i < n) {
				uint32_t p = list ? list[i] : i;
				act |= 1u << r;
				key[r] = canonical_key<KW>(R.W, R.RC, p, P.k, Lp, P.mask_hi, P.mask_lo);
				const uint64_t h = key_hash<KW>(key[r]);
				// the membership filter first: windows that end up here are mostly not in the draft at all
				if (P.bloom && !bloom_maybe(P.bloom, bloom_probe(h, P.bloom_words))) {
					act &= ~(1u << r);
					continue;
				}
				load_slot(P.table, hash_to_slot(h, P.nslots), hi[r], lo[r], val[r], pi);
			}
		}
		uint32_t hit[kProbeBatch];
#pragma unroll
		for (int r = 0; r < kProbeBatch; ++r) {
			hit[r] = 0;
			if (act & (1u << r)) {
				bool found = slot_matches<KW>(hi[r], lo[r], key[r]);
				bool empty = slot_empty<KW>(hi[r], lo[r]);
				if (!found && !empty) {
					// rare: the home slot holds another key -- walk the probe sequence
					uint64_t slot = hash_to_slot(key_hash<KW>(key[r]), P.nslots);
					do {
						slot = slot + 1 == P.nslots ? 0 : slot + 1;
						load_slot(P.table, slot, hi[r], lo[r], val[r], pi);
						found = slot_matches<KW>(hi[r], lo[r], key[r]);
						empty = slot_empty<KW>(hi[r], lo[r]);
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

// classify windows [0, nw) of a region as valid / invalid (kv / ki) without resolving them
__device__ __forceinline__ void
warp_count_windows(const WarpRegion& R, uint32_t nw, uint32_t k, uint32_t lane, uint16_t* list, uint32_t& n_list, LaneStats& st)
{
	n_list = 0;
	for (uint32_t pb = 0; pb < nw; pb += 32) {
		const uint32_t p = pb + lane;
		bool need = false;
		if (p < nw) {
			if (window_invalid(R.INV, p, k)) {
				st.ki++;
			} else {
				st.kv++;
				need = true;
			}
		}
		const uint32_t m = __ballot_sync(0xFFFFFFFFu, need);
		if (need)
			list[n_list + __popc(m & ((1u << lane) - 1u))] = (uint16_t)p;
		n_list += __popc(m);
	}
	__syncwarp();
}

// bestContig's window loop for a read longer than one region: repack chunk by chunk (cold path)
// (cold path, not inlined: it gets COPIES of the parameter block, the vote table and the counters so
// that the hot path's objects never have their address taken and stay in registers / the constant bank)
struct LongResult
{
	Track tr;
	LaneStats st;
};
template <int KW>
__device__ __noinline__ LongResult
warp_windows_long(WarpRegion& R, uint16_t* list, const char* src, uint32_t total, MapParams P, uint32_t lane, Track tr, LaneStats st)
{
	const uint32_t cw = kRegionBases - P.k + 1;
	for (uint32_t c0 = 0; c0 < total; c0 += cw) {
		uint32_t nwc = min(cw, total - c0);
		uint32_t Lc = nwc + P.k - 1;
		uint32_t a, b, n_list;
		__syncwarp();
		warp_pack(R, src + c0, Lc, lane, a, b);
		warp_count_windows(R, nwc, P.k, lane, list, n_list, st);
		warp_probe_windows<KW>(R, Lc, list, n_list, P, lane, tr, st);
	}
	return LongResult{tr, st};
}

// 16 mismatch bits (LSB-first by base) between a packed stream word and the contig text at
// global base coordinate gb (may be out of range: then every base mismatches)
__device__ __forceinline__ uint32_t mismatch16(uint32_t sword, const uint32_t* T, int64_t gb, int64_t n_bases)
{
	if (gb < 0 || gb + 16 > n_bases)
		return 0xFFFFu;
	const uint64_t wi = (uint64_t)gb >> 4;
	const uint32_t sh = ((uint32_t)gb & 15u) * 2u;
	const uint32_t t0 = __ldg(T + wi), t1 = __ldg(T + wi + 1);
	const uint32_t t = __funnelshift_l(t1, t0, sh);
	const uint32_t x = sword ^ t;
	uint32_t mm = (x | (x >> 1)) & 0x55555555u; // bit (30-2i) = base i differs
	uint32_t y = __brev(mm) >> 1;               // bit 2i = base i differs
	y = (y | (y >> 1)) & 0x33333333u;
	y = (y | (y >> 2)) & 0x0F0F0F0Fu;
	y = (y | (y >> 4)) & 0x00FF00FFu;
	y = (y | (y >> 8)) & 0x0000FFFFu;
	return y;
}

struct Extension
{
	bool have;     // a seed hit established a diagonal
	bool same;     // read and contig on the same strand
	int64_t D;     // contig coordinate of stream coordinate 0
	int64_t lo, hi; // the seed's contig end covers [lo, hi)
	uint32_t c_end; // its contig-end record
	uint32_t off;  // stream offset of read coordinate 0 in the RC stream (Lp - L)
};

// bestContig (Arcs.cpp:939-1014) for one prepacked mate.  Seed-and-extend: the seed's slot
// says where its k-mer sits in the packed contig text; the read is compared with the text
// along that diagonal, and every window whose k bases all match a window that mapKmers
// inserted is resolved without touching the table (same key => same value; the value is the
// contig end if the TUNIQ bit is set, else 0).  Everything else is probed as before.
// `clean` = the mate has no invalid character at all (then no window is invalid).
template <int KW>
__device__ __forceinline__ void
warp_resolve_read(const WarpRegion& R, uint32_t* M, uint16_t* list, uint32_t L, uint32_t total, bool clean, const Extension& E,
    const MapParams& P, uint32_t lane, Track& tr, LaneStats& st)
{
	const uint32_t nwords = (L + 15) >> 4;
	uint32_t n_list = 0;
	if (E.have) {
		// mismatch mask in stream coordinates; bases outside the read are forced to "match"
		const uint32_t* S = E.same ? R.W : R.RC;
		const uint32_t q0 = E.same ? 0u : E.off; // stream coordinate of the first read base / window
		uint32_t m16 = 0;
		if (lane < nwords) {
			m16 = mismatch16(S[lane], P.ct_T, E.D + 16 * (int64_t)lane, (int64_t)P.ct_n_bases);
			const uint32_t b0 = 16 * lane; // keep only stream coordinates in [q0, q0 + L)
			const uint32_t lo = q0 > b0 ? min(q0 - b0, 16u) : 0u;
			const uint32_t hi = q0 + L > b0 ? min(q0 + L - b0, 16u) : 0u;
			m16 &= (hi > lo) ? (((1u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
		}
		const uint32_t any_mm = __ballot_sync(0xFFFFFFFFu, m16 != 0);
		const int64_t gw0 = E.D + q0; // contig coordinate of stream window q0
		const bool inside = gw0 >= E.lo && gw0 + (int64_t)(total - 1) + (int64_t)P.k <= E.hi;
		if (any_mm == 0 && inside && clean) {
			// ---- fast path: the whole mate equals the contig text: a window is found iff it was
			// inserted, recorded iff its key is unique; word-parallel over the two bit masks
			const uint64_t w0 = (uint64_t)gw0 >> 5;
			const uint32_t sh = (uint32_t)gw0 & 31u;
			const uint32_t nw_words = (sh + total + 31) >> 5; // <= 17
			uint32_t ins = 0, uq = 0, range = 0;
			if (lane < nw_words) {
				const uint32_t first = lane == 0 ? sh : 0u;
				const uint32_t last = min(32u, sh + total - 32u * lane); // exclusive
				range = (last >= 32 ? 0xFFFFFFFFu : ((1u << last) - 1u)) & ~((1u << first) - 1u);
				ins = __ldg(P.ct_TINS + w0 + lane) & range;
				uq = __ldg(P.ct_TUNIQ + w0 + lane) & ins;
			}
			st.kv += lane == 0 ? total : 0u;
			st.found += __popc(ins);
			st.rec += __popc(uq);
			st.dups += __popc(ins & ~uq);
			const uint32_t cnt = warp_sum(__popc(uq));
			if (cnt)
				track_add(tr, lane, E.c_end, cnt);
			// windows that were not inserted (skipped by the N walk): rare, probe them
			uint32_t miss = range & ~ins;
			const uint32_t any_miss = __ballot_sync(0xFFFFFFFFu, miss != 0);
			if (any_miss) {
				for (uint32_t src = 0; src < nw_words; ++src) {
					uint32_t mw = __shfl_sync(0xFFFFFFFFu, miss, src);
					if (lane == 0) {
						while (mw) {
							const uint32_t bit = __ffs(mw) - 1;
							mw &= mw - 1;
							const uint32_t q = (uint32_t)(((w0 + src) << 5) + bit - (uint64_t)E.D);
							list[n_list++] = (uint16_t)(E.same ? q : (L - P.k) - (q - E.off));
						}
					}
					n_list = __shfl_sync(0xFFFFFFFFu, n_list, 0);
				}
				__syncwarp();
				warp_probe_windows<KW>(R, L, list, n_list, P, lane, tr, st);
			}
			return;
		}
		const uint32_t up = __shfl_down_sync(0xFFFFFFFFu, m16, 1);
		if ((lane & 1) == 0)
			M[lane >> 1] = m16 | (up << 16);
		// the inserted / unique window masks along the diagonal, loaded once (one word per lane) so that
		// the window loop below touches shared memory only; words outside the text read as 0 = unresolved
		{
			const int64_t wfirst = gw0 >> 5; // arithmetic shift: may be negative
			const int64_t wi = wfirst + (int64_t)lane;
			const bool in = lane < (uint32_t)kMaskWords && wi >= 0 && wi < (int64_t)((P.ct_n_bases + 31) >> 5);
			if (lane < (uint32_t)kMaskWords) {
				M[kRegionInvWords + lane] = in ? __ldg(P.ct_TINS + wi) : 0u;
				M[kRegionInvWords + kMaskWords + lane] = in ? __ldg(P.ct_TUNIQ + wi) : 0u;
			}
		}
		__syncwarp();
	}
	uint32_t ext_uniq = 0;
	const uint32_t* TI = M + kRegionInvWords;
	const uint32_t* TU = TI + kMaskWords;
	const int64_t gw_first_word = E.have ? ((E.D + (int64_t)(E.same ? 0u : E.off)) >> 5) : 0;
	for (uint32_t pb = 0; pb < total; pb += 32) {
		const uint32_t p = pb + lane;
		bool need = false;
		if (p < total) {
			if (!clean && window_invalid(R.INV, p, P.k)) {
				st.ki++;
			} else {
				st.kv++;
				need = true;
				if (E.have) {
					const uint32_t q = E.same ? p : (L - P.k - p) + E.off;
					if (!window_invalid(M, q, P.k)) {
						const int64_t gw = E.D + q;
						if (gw >= E.lo && gw + (int64_t)P.k <= E.hi) {
							const uint32_t wi = (uint32_t)((gw >> 5) - gw_first_word), bit = (uint32_t)gw & 31u;
							if ((TI[wi] >> bit) & 1u) {
								need = false;
								st.found++;
								if ((TU[wi] >> bit) & 1u) {
									st.rec++;
									ext_uniq++;
								} else {
									st.dups++;
								}
							}
						}
					}
				}
			}
		}
		const uint32_t m = __ballot_sync(0xFFFFFFFFu, need);
		if (need)
			list[n_list + __popc(m & ((1u << lane) - 1u))] = (uint16_t)p;
		n_list += __popc(m);
	}
	__syncwarp();
	if (E.have) {
		const uint32_t cnt = warp_sum(ext_uniq);
		if (cnt)
			track_add(tr, lane, E.c_end, cnt);
	}
	warp_probe_windows<KW>(R, L, list, n_list, P, lane, tr, st);
}

// warp-uniform per-pair / per-read counters
struct PairCounters
{
	uint32_t pass, fail, stored, invalid, nogood;
	bool overflow;
};

// argmax over the vote table (ties -> smallest contig end) + Jaccard gate (Arcs.cpp:996-1012).
// Returns the read's contig end or 0; *passed tells which Jaccard counter to bump.
__device__ __forceinline__ void warp_best(const Track& tr, uint32_t lane, uint32_t& best_cnt, uint32_t& best_c)
{
	const uint32_t mycnt = lane < tr.n ? tr.cnt : 0;
	best_cnt = __reduce_max_sync(0xFFFFFFFFu, mycnt);
	const uint32_t cand = (best_cnt && lane < tr.n && tr.cnt == best_cnt) ? tr.c : 0xFFFFFFFFu;
	best_c = __reduce_min_sync(0xFFFFFFFFu, cand);
}

__device__ __forceinline__ uint32_t
warp_decide(uint32_t best_cnt, uint32_t best_c, uint32_t total, const MapParams& P, bool* passed)
{
	bool ok;
	if (best_cnt == 0)
		ok = 0.0 > P.j_index;
	else if (total <= (uint32_t)kRegionBases)
		ok = best_cnt >= __ldg(P.jmin + total);
	else
		ok = (double)best_cnt / (double)total > P.j_index;
	*passed = ok;
	return (ok && best_cnt) ? best_c : 0u;
}

__device__ __forceinline__ void prefetch_l1(const void* p)
{
	asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// Issues L1 prefetches for everything the general path will touch for this work item: the ASCII
// bases of the pair, and per unfinished mate with a seed the contig text, the two window masks and
// the contig-end metadata along the seed's diagonal.  One line per lane.
__device__ __forceinline__ void prefetch_item(const WorkRecord& rec, const MapParams& P, uint32_t lane)
{
	const uint32_t o0 = rec.off[0], o2 = rec.off[2];
	if (lane < 6) {
		const char* p = P.bases + o0 + 128u * lane;
		if (p < P.bases + o2 + 127)
			prefetch_l1(p);
		return;
	}
	const uint32_t rd = (lane - 6) / 8, what = (lane - 6) % 8;
	if (rd > 1 || rec.state[rd] != kMateSlow || !(rec.meta[rd] & 1u))
		return;
	const uint64_t pi = ((uint64_t)rec.seed_hi[rd] << 32) | rec.seed_lo[rd];
	const uint64_t g = pi & kPosMask;
	const uint32_t eidx = (uint32_t)(pi >> 40);
	// the diagonal starts within one read length of g on either side: cover [g - 256, g + 256)
	const uint64_t g_lo = g > 256 ? g - 256 : 0;
	switch (what) {
	case 0: prefetch_l1(P.ct_T + (g_lo >> 4)); break;
	case 1: prefetch_l1(P.ct_T + (g_lo >> 4) + 32); break;
	case 2: prefetch_l1(P.ct_TINS + (g_lo >> 5)); break;
	case 3: prefetch_l1(P.ct_TINS + (g_lo >> 5) + 16 < P.ct_TINS + (P.ct_n_bases >> 5) ? P.ct_TINS + (g_lo >> 5) + 16 : P.ct_TINS); break;
	case 4: prefetch_l1(P.ct_TUNIQ + (g_lo >> 5)); break;
	case 5: prefetch_l1(P.ct_end_g0 + eidx); break;
	case 6: prefetch_l1(P.ct_end_len + eidx); break;
	default: prefetch_l1(P.ct_end_cr + eidx); break;
	}
}

// One read pair handled by the whole warp (any read length).  Generic path.
// state0/state1: kMateUnknown for a pair nobody has looked at (validity is checked here), else
// per mate either kMateSlow (resolve it here) or the contig end already decided elsewhere
// (then that mate is not touched and none of its counters are bumped).
#ifdef ARKS_PAIR_NOINLINE
#define ARKS_PAIR_ATTR __noinline__
#else
#define ARKS_PAIR_ATTR __forceinline__
#endif
template <int KW>
__device__ ARKS_PAIR_ATTR uint32_t
warp_process_pair(const WorkRecord& rec, const MapParams& P, WarpRegion* R, uint32_t* M, uint16_t* list, uint32_t lane, LaneStats& st,
    PairCounters& pc)
{
	const uint32_t pair = rec.pair, state0 = rec.state[0], state1 = rec.state[1];
	const bool fresh = state0 == kMateUnknown;
	const bool todo0 = fresh || state0 == kMateSlow, todo1 = fresh || state1 == kMateSlow;
	const uint32_t o0 = rec.off[0], o1 = rec.off[1], o2 = rec.off[2];
	const uint32_t l1 = o1 - o0, l2 = o2 - o1;
	const bool shortpair = l1 <= (uint32_t)kRegionBases && l2 <= (uint32_t)kRegionBases;
	uint32_t nn1, no1, nn2, no2;
	__syncwarp();
	if (l1 <= 256u && l2 <= 256u) {
		warp_pack2(R, P.bases + o0, l1, P.bases + o1, l2, lane, nn1, no1, nn2, no2);
	} else if (shortpair) {
		warp_pack(R[0], P.bases + o0, l1, lane, nn1, no1);
		warp_pack(R[1], P.bases + o1, l2, lane, nn2, no2);
	} else {
		warp_classify(P.bases + o0, l1, lane, nn1, no1);
		warp_classify(P.bases + o1, l2, lane, nn2, no2);
	}
	uint32_t c[2] = {fresh ? 0u : state0, fresh ? 0u : state1};
	if (!fresh || (read_ok(nn1, no1, l1, P.nmax) && read_ok(nn2, no2, l2, P.nmax))) {
		// ---- seeds of both mates in one round: lanes [0,kSeeds) mate 0, [kSeeds,2kSeeds) mate 1
		uint64_t seed_pos = 0;
		uint32_t seed_p = 0;
		bool seed_found = false, seed_fc = false;
		if (fresh && shortpair && P.use_extension) {
			const uint32_t rd = lane / kSeeds, sidx = lane % kSeeds;
			const uint32_t len = rd ? l2 : l1;
			if (lane < 2 * kSeeds && len >= P.k) {
				const uint32_t total = len - P.k + 1;
				seed_p = kSeeds > 1 ? (uint32_t)(((uint64_t)(total - 1) * sidx) / (kSeeds - 1)) : 0;
				const WarpRegion& Rr = R[rd];
				if (!window_invalid(Rr.INV, seed_p, P.k)) {
					const uint32_t Lp = ((len + 15) >> 4) << 4;
					Key128 key = canonical_key<KW>(Rr.W, Rr.RC, seed_p, P.k, Lp, P.mask_hi, P.mask_lo, &seed_fc);
					uint64_t slot = hash_to_slot(key_hash<KW>(key), P.nslots);
					while (true) {
						uint64_t hi, lo;
						uint32_t val;
						load_slot(P.table, slot, hi, lo, val, seed_pos);
						if (slot_matches<KW>(hi, lo, key)) {
							seed_found = true;
							break;
						}
						if (slot_empty<KW>(hi, lo))
							break;
						slot = slot + 1 == P.nslots ? 0 : slot + 1;
					}
				}
			}
		}
		const uint32_t seed_votes = __ballot_sync(0xFFFFFFFFu, seed_found);
#pragma unroll 1
		for (int rd = 0; rd < 2; ++rd) {
			// bestContig (Arcs.cpp:939-1014) for mate rd
			if (!(rd ? todo1 : todo0))
				continue;
			const uint32_t len = rd ? l2 : l1;
			const uint32_t total = len >= P.k ? len - P.k + 1 : 0;
			// votes of the read; repeated in 2, 4, ... passes if it votes for more contig ends than a warp tracks
			uint32_t best_cnt = 0, best_c = 0xFFFFFFFFu;
			const LaneStats st0 = st;
			uint32_t parts = 1;
#pragma unroll 1
			for (uint32_t part = 0; part < parts; ++part) {
			st = st0; // every pass visits every window: the counters of the last pass are the read's
			Track tr{0, 0, 0, false, part, parts};
			if (total) {
				if (shortpair) {
					Extension E{};
					uint64_t pi;
					uint32_t sp;
					bool r_fc;
					if (fresh) {
						const uint32_t mine = (seed_votes >> (rd * kSeeds)) & ((1u << kSeeds) - 1u);
						E.have = mine != 0;
						const int src = E.have ? rd * kSeeds + __ffs(mine) - 1 : 0;
						pi = __shfl_sync(0xFFFFFFFFu, seed_pos, src);
						sp = __shfl_sync(0xFFFFFFFFu, seed_p, src);
						r_fc = __shfl_sync(0xFFFFFFFFu, (int)seed_fc, src);
					} else { // the group kernel already probed this mate's seeds
						const uint32_t meta = rec.meta[rd];
						E.have = (meta & 1u) != 0;
						pi = ((uint64_t)rec.seed_hi[rd] << 32) | rec.seed_lo[rd];
						sp = meta >> 8;
						r_fc = (meta & 2u) != 0;
					}
					if (E.have) {
						const bool c_fc = (pi >> kPosBits) & 1ull;
						const uint64_t g = pi & kPosMask;
						const uint32_t eidx = (uint32_t)(pi >> 40);
						E.same = r_fc == c_fc;
						E.off = (((len + 15) >> 4) << 4) - len;
						E.D = E.same ? (int64_t)g - (int64_t)sp : (int64_t)g - (int64_t)((len - P.k - sp) + E.off);
						E.lo = (int64_t)__ldg(P.ct_end_g0 + eidx);
						E.hi = E.lo + (int64_t)__ldg(P.ct_end_len + eidx);
						E.c_end = __ldg(P.ct_end_cr + eidx);
					}
					const bool clean = (rd ? (nn2 | no2) : (nn1 | no1)) == 0;
					warp_resolve_read<KW>(R[rd], M, list, len, total, clean, E, P, lane, tr, st);
				} else {
					const LongResult lr = warp_windows_long<KW>(R[0], list, P.bases + (rd ? o1 : o0), total, P, lane, tr, st);
					tr = lr.tr;
					st = lr.st;
				}
			}
			if (tr.overflow) { // warp-uniform
				if (parts >= 65536u) {
					pc.overflow = true; // more distinct contig ends than windows: cannot happen
				} else {
					parts *= 2;
					part = 0xFFFFFFFFu; // start over with twice as many passes
					best_cnt = 0;
					best_c = 0xFFFFFFFFu;
					continue;
				}
			}
			uint32_t cnt_p, c_p;
			warp_best(tr, lane, cnt_p, c_p);
			if (cnt_p > best_cnt || (cnt_p == best_cnt && cnt_p && c_p < best_c)) {
				best_cnt = cnt_p;
				best_c = c_p;
			}
			}
			bool passed;
			c[rd] = warp_decide(best_cnt, best_c, total, P, &passed);
			if (passed)
				pc.pass++;
			else
				pc.fail++;
		}
	} else {
		pc.invalid++;
	}
	uint32_t out = 0;
	if (c[0] != 0 && c[0] == c[1]) {
		pc.stored++;
		out = c[0];
		if (lane == 0) {
			uint32_t cc = (P.remap && out < P.n_remap) ? P.remap[out] : out;
			imap_add(P.imap, P.imap_mask, P.imap_count, P.barcode_id[pair], (cc - 1) >> 1, (cc & 1u), (cc & 1u) ^ 1u);
		}
	} else {
		pc.nogood++;
	}
	if (P.conreci_out && lane == 0)
		P.conreci_out[pair] = (int32_t)out;
	return out;
}

__device__ __forceinline__ void
flush_counters(const MapParams& P, uint32_t lane, const LaneStats& st, uint32_t pass, uint32_t fail, uint32_t stored, uint32_t invalid,
    uint32_t nogood, bool overflow)
{
	// every counter is a lane-private partial sum
	const uint32_t kv = warp_sum(st.kv), ki = warp_sum(st.ki), fo = warp_sum(st.found), re = warp_sum(st.rec), du = warp_sum(st.dups);
	pass = warp_sum(pass);
	fail = warp_sum(fail);
	stored = warp_sum(stored);
	invalid = warp_sum(invalid);
	nogood = warp_sum(nogood);
	const uint32_t ov = __ballot_sync(0xFFFFFFFFu, overflow);
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
		if (ov) atomicAdd(&ctr->overflow, 1ull);
	}
}

// ---- warp-per-pair kernel (generic; kept for A/B runs: ARKS_MAP_MODE=pair) ------------------
template <int KW>
__global__ void __launch_bounds__(kMapThreads, kMapMinBlocks) map_pairs_kernel(MapParams P)
{
	__shared__ WarpRegion regions[kMapWarps][2];
	__shared__ uint32_t Ms[kMapWarps][kMWords];
	__shared__ uint16_t lists[kMapWarps][kRegionBases];
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warp = threadIdx.x >> 5;
	LaneStats st{0, 0, 0, 0, 0};
	PairCounters pc{0, 0, 0, 0, 0, false};
	const uint32_t nwarps = gridDim.x * kMapWarps;
#pragma unroll 1
	for (uint32_t pair = blockIdx.x * kMapWarps + warp; pair < P.n_pairs; pair += nwarps) {
		WorkRecord rec{};
		rec.pair = pair;
		rec.off[0] = P.read_off[2 * pair];
		rec.off[1] = P.read_off[2 * pair + 1];
		rec.off[2] = P.read_off[2 * pair + 2];
		rec.state[0] = rec.state[1] = kMateUnknown;
		warp_process_pair<KW>(rec, P, regions[warp], Ms[warp], lists[warp], lane, st, pc);
	}
	const bool l0 = lane == 0; // pair counters are warp-uniform: count them once
	flush_counters(P, lane, st, l0 ? pc.pass : 0, l0 ? pc.fail : 0, l0 ? pc.stored : 0, l0 ? pc.invalid : 0, l0 ? pc.nogood : 0,
	    pc.overflow);
}

// ---- warp-per-16-pairs kernel -----------------------------------------------------------------
// A warp takes a group of 16 consecutive pairs = 32 reads.
//   stage 1  all lanes pack the 32 reads to 2 bits (flat over the reads' 16-base words, so no lane
//            idles), invalid-base masks alongside;
//   stage 2  ONE LANE PER READ: seed probes (two in flight per lane, 64 per warp), the
//            comparison of the whole read with the packed contig text along the seed's diagonal
//            (a mismatch bit per base), and the word-parallel window classification: a window whose
//            k bases are valid and equal the text is found iff the text window was inserted and
//            recorded iff its key is unique (two bit masks of the contig text); a window that
//            overlaps a mismatch, or matches a text window that was not inserted, has to be looked
//            up.  A read that equals the contig text -- the common case -- is finished here;
//   stage 2b ALL LANES, FLAT over the windows the 32 reads still have to look up: key, hash, the
//            L2-resident membership filter, and the table only on a filter positive (these windows
//            carry a sequencing error: almost none of them is in the draft);
//   stage 3  one lane per pair: if both mates are finished, the pair rule (Arcs.cpp:1280) and
//            the barcode tally; otherwise the pair goes to a work list with the state of each
//            mate, and map_slow_kernel resolves the unfinished mates (no seed hit, not inside one
//            contig end, votes for a second contig end, reads longer than kGroupReadBases) with
//            the general warp-per-pair path.
#ifndef ARKS_GROUP_WARPS
#define ARKS_GROUP_WARPS 8
#endif
#ifndef ARKS_GROUP_MIN_BLOCKS
#define ARKS_GROUP_MIN_BLOCKS 3
#endif
#ifndef ARKS_GROUP_SEEDS
#define ARKS_GROUP_SEEDS 4
#endif
#ifndef ARKS_GROUP_LIST
#define ARKS_GROUP_LIST 512
#endif
constexpr int kGroupWarps = ARKS_GROUP_WARPS;
constexpr int kGroupThreads = kGroupWarps * 32;
constexpr int kGroupMinBlocks = ARKS_GROUP_MIN_BLOCKS;
constexpr int kGroupSeeds = ARKS_GROUP_SEEDS;  // seed windows tried per read (two at a time)
constexpr int kGroupListCap = ARKS_GROUP_LIST; // lookups staged per round of stage 2b
constexpr int kGroupPairs = 16;
constexpr int kGroupReadBases = 256;
constexpr int kGroupWStride = kGroupReadBases / 16 + 5; // 21 words: odd stride, +4 slack for extraction
constexpr int kGroupIStride = kGroupReadBases / 32 + 3; // 11 words
constexpr int kGroupMStride = kGroupReadBases / 32 + 1; // 9 words: one mask bit per base/window + a zero pad word

struct GroupSmem
{
	uint32_t W[32][kGroupWStride];
	// the invalid-base / mismatch masks are dead once the windows are classified; the lookup lists of
	// stage 2b live in the same bytes
	union Scratch
	{
		struct Masks
		{
			uint32_t INV[32][kGroupIStride];
			uint32_t BM[32][kGroupMStride]; // per read: mismatch/invalid base mask, then "window overlaps one" (text orientation)
		} m;
		struct Lists
		{
			uint16_t list[kGroupListCap]; // staged lookups: read << 11 | first window (text orientation) << 3 | windows - 1
			uint16_t pos[256];            // filter positives of one round: read << 8 | read window
			uint32_t npos;
		} l;
	} s;
	uint32_t PM[32][kGroupMStride]; // per read: invalid base mask, then "window overlaps one", then the lookup mask
	uint32_t woff[33];
	uint32_t nbad[32];   // per read: N count | other-invalid count << 16
	uint32_t rinfo[32];  // per read: number of windows | same-strand flag << 16
	uint32_t rcend[32];  // per read: the seed's contig end
	uint32_t rcend2[32]; // per read: a second contig end its lookups voted for (0: none yet)
	uint32_t pcnt[32];   // per read: lookups found | lookups recorded << 16
	uint32_t pcnt2[32];  // per read: lookups recorded for rcend2 | (a third contig end turned up) << 31
};
static_assert(sizeof(GroupSmem::Scratch::Lists) <= sizeof(GroupSmem::Scratch::Masks), "lookup lists must fit in the mask rows");

#ifndef ARKS_MM_UNROLL
#define ARKS_MM_UNROLL 1
#endif
constexpr int kMismatchUnroll = ARKS_MM_UNROLL; // unroll factor of the read-vs-text comparison loop
#ifndef ARKS_CHUNK
#define ARKS_CHUNK 8
#endif
constexpr uint32_t kChunkWindows = ARKS_CHUNK; // consecutive windows (<= 8) one lane looks up with a rolling key
#ifndef ARKS_LOOKUP_BATCH
#define ARKS_LOOKUP_BATCH 1
#endif
constexpr int kLookupBatch = ARKS_LOOKUP_BATCH; // filter loads in flight per lane

// bits b of word w with lo <= 32 w + b < hi
__device__ __forceinline__ uint32_t word_range_mask(uint32_t lo, uint32_t hi, uint32_t w)
{
	const uint32_t w0 = 32u * w;
	const uint32_t a = lo > w0 ? min(lo - w0, 32u) : 0u;
	const uint32_t b = hi > w0 ? min(hi - w0, 32u) : 0u;
	if (b <= a)
		return 0u;
	return (b >= 32u ? 0xFFFFFFFFu : ((1u << b) - 1u)) & ~((1u << a) - 1u);
}

// number of chunks of at most kChunkWindows consecutive set bits a mask word splits into
__device__ __forceinline__ uint32_t count_chunks(uint32_t m)
{
	uint32_t n = 0;
	while (m) {
		const uint32_t s = __ffs(m) - 1;
		const uint32_t t = m >> s;
		const uint32_t run = t == 0xFFFFFFFFu ? 32u : (uint32_t)__ffs(~t) - 1u;
		const uint32_t c = min(run, kChunkWindows);
		m &= ~((c >= 32u ? 0xFFFFFFFFu : ((1u << c) - 1u)) << s);
		n++;
	}
	return n;
}

// seed windows in the order they are tried: both ends first, then the middle, then the quarters
__device__ __forceinline__ uint32_t seed_window(uint32_t sidx, uint32_t total)
{
	const uint32_t last = total - 1;
	switch (sidx) {
	case 0: return 0;
	case 1: return last;
	case 2: return last >> 1;
	case 3: return last >> 2;
	default: return (3 * last) >> 2;
	}
}

// bits [s, s + 32) of a LSB-first mask row restricted to [0, len); s may be negative.  The row has
// a readable word after the one that holds bit len - 1.
__device__ __forceinline__ uint32_t mask_bits_at(const uint32_t* row, int s, int len)
{
	if (s >= len || s + 32 <= 0)
		return 0u;
	const int lo = s > 0 ? s : 0;
	uint32_t v = __funnelshift_r(row[lo >> 5], row[(lo >> 5) + 1], (uint32_t)lo & 31u);
	const int nvalid = len - lo;
	if (nvalid < 32)
		v &= (1u << nvalid) - 1u;
	if (s < 0)
		v <<= (uint32_t)(-s);
	return v;
}

// X[u] |= X[u + a] over a row of nmw words (words past the row read as 0); in place, ascending
__device__ __forceinline__ void or_shifted_row(uint32_t* X, uint32_t nmw, uint32_t a)
{
	const uint32_t ws = a >> 5, bs = a & 31u;
	for (uint32_t i = 0; i < nmw; ++i) {
		const uint32_t i0 = i + ws;
		const uint32_t lo = i0 < nmw ? X[i0] : 0u;
		const uint32_t hi = i0 + 1 < nmw ? X[i0 + 1] : 0u;
		X[i] |= __funnelshift_r(lo, hi, bs);
	}
}

// base mask -> window mask: X[u] = OR of X[u .. u + k - 1]  (log2(k) doubling passes)
__device__ __forceinline__ void dilate_row(uint32_t* X, uint32_t nmw, uint32_t k)
{
	uint32_t span = 1;
	while (span * 2 <= k) {
		or_shifted_row(X, nmw, span);
		span *= 2;
	}
	if (k > span)
		or_shifted_row(X, nmw, k - span);
}

// The lane-per-read kernel is instruction-cache bound as soon as its hot loop outgrows ~2.5k SASS
// instructions (profiles/: stall_no_inst 4 % -> 36 % between two builds that differ in code size only),
// so everything that is needed at several places or only rarely is a real function call:

// canonical key of read window p (forward packing in W) and its slot hash
struct KeyHash
{
	Key128 key;
	uint64_t hash;
	uint32_t fwd_is_canonical;
};
template <int KW>
__device__ __forceinline__ KeyHash window_key_hash(const uint32_t* W, uint32_t p, uint32_t k, uint64_t mask_hi, uint64_t mask_lo)
{
	KeyHash r;
	bool fc;
	const Key128 f = extract_window<KW>(W, p, mask_hi, mask_lo);
	r.key = canonical_from_forward<KW>(f, k, &fc);
	r.hash = key_hash<KW>(r.key);
	r.fwd_is_canonical = fc;
	return r;
}

struct ChainResult
{
	uint64_t posinfo;
	uint32_t val;
	uint32_t found;
};

// the home slot held another key: walk the probe sequence (rare at load 0.5)
template <int KW>
__device__ __noinline__ ChainResult slot_chain_find(const uint8_t* table, uint64_t nslots, uint64_t slot, uint64_t key_hi, uint64_t key_lo)
{
	const Key128 key{key_hi, key_lo};
	ChainResult r;
	while (true) {
		slot = slot + 1 == nslots ? 0 : slot + 1;
		uint64_t hi, lo;
		load_slot(table, slot, hi, lo, r.val, r.posinfo);
		r.found = slot_matches<KW>(hi, lo, key);
		if (r.found || slot_empty<KW>(hi, lo))
			return r;
	}
}

__device__ __noinline__ void dilate_row_call(uint32_t* X, uint32_t nmw, uint32_t k)
{
	dilate_row(X, nmw, k);
}

// the group a warp takes after `group`: the next one nobody has taken yet (a ticket counter behind the work-list
// counter), which evens out groups of different cost -- +5 % on configs[1], +3 % on configs[2] against a fixed stride
// (-DARKS_STATIC_GROUPS, kept for A/B runs)
__device__ __forceinline__ uint32_t next_group(const MapParams& P, uint32_t lane, uint32_t nwarps, uint32_t group)
{
#ifndef ARKS_STATIC_GROUPS
	uint32_t t = 0;
	if (lane == 0)
		t = atomicAdd(P.work_count + 1, 1u);
	return nwarps + __shfl_sync(0xFFFFFFFFu, t, 0);
#else
	return group + nwarps;
#endif
}

template <int KW>
__global__ void __launch_bounds__(kGroupThreads, kGroupMinBlocks) map_groups_kernel(MapParams P)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	GroupSmem& G = reinterpret_cast<GroupSmem*>(smem_raw)[threadIdx.x >> 5];
	const uint32_t lane = threadIdx.x & 31;
	LaneStats st{0, 0, 0, 0, 0};
	uint32_t pass = 0, fail = 0, stored = 0, invalid = 0, nogood = 0;
	const uint32_t n_groups = (P.n_pairs + kGroupPairs - 1) / kGroupPairs;
	const uint32_t nwarps = gridDim.x * kGroupWarps;
#pragma unroll 1
	for (uint32_t group = blockIdx.x * kGroupWarps + (threadIdx.x >> 5); group < n_groups; group = next_group(P, lane, nwarps, group)) {
		const uint32_t pair0 = group * kGroupPairs;
		const uint32_t my_pair = pair0 + (lane >> 1);
		const bool exists = my_pair < P.n_pairs;
		// ---- read lane's extent
		uint32_t off = 0, len = 0;
		if (exists) {
			off = P.read_off[2 * pair0 + lane];
			len = P.read_off[2 * pair0 + lane + 1] - off;
		}
		const bool is_long = len > (uint32_t)kGroupReadBases;
		const uint32_t long_pairs = __ballot_sync(0xFFFFFFFFu, is_long);
		const bool pair_long = ((long_pairs >> (lane & ~1u)) & 3u) != 0;
		const uint32_t nwords = (exists && !pair_long) ? (len + 15) >> 4 : 0;
		// ---- stage 1: flat packing
		uint32_t incl = nwords;
#pragma unroll 1
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
			if (lane >= (uint32_t)o)
				incl += t;
		}
		__syncwarp();
		G.woff[lane + 1] = incl;
		if (lane == 0)
			G.woff[0] = 0;
		G.nbad[lane] = 0;
		const uint32_t total_words = __shfl_sync(0xFFFFFFFFu, incl, 31);
		__syncwarp();
#ifndef ARKS_NO_FAST_DIV
		// all 32 reads need the same number of words (the usual case): read index by multiplication
		const uint32_t nw0 = __shfl_sync(0xFFFFFFFFu, nwords, 0);
		const bool uniform = nw0 != 0 && __all_sync(0xFFFFFFFFu, nwords == nw0);
		const uint32_t magic = nw0 ? 0xFFFFFFFFu / nw0 + 1u : 0u;
#endif
		for (uint32_t f = lane; f < total_words; f += 32) {
			// read r with woff[r] <= f < woff[r+1]
			uint32_t r = 0;
#ifndef ARKS_NO_FAST_DIV
			if (uniform) {
				r = __umulhi(f, magic);
			} else
#endif
			{
#pragma unroll
				for (int step = 16; step > 0; step >>= 1)
					if (G.woff[r + step] <= f)
						r += step;
			}
			const uint32_t j = f - G.woff[r];
			const uint32_t roff = P.read_off[2 * pair0 + r];
			const uint32_t rlen = P.read_off[2 * pair0 + r + 1] - roff;
			uint32_t inv16, nn, no;
			G.W[r][j] = pack_group_compact(P.bases + roff, rlen, j, &inv16, &nn, &no);
			reinterpret_cast<uint16_t*>(G.s.m.INV[r])[j] = (uint16_t)inv16;
			if (inv16)
				atomicAdd(&G.nbad[r], nn | (no << 16));
		}
		__syncwarp();
		// ---- stage 2: one lane per read
		const uint32_t bad = G.nbad[lane];
		const uint32_t n_n = bad & 0xFFFFu, n_other = bad >> 16;
		const bool clean = bad == 0;
		// checkReadSequence (Arcs.cpp:366-389) through the exact integer table (reads here are <= kGroupReadBases)
		const bool ok = exists && !pair_long && n_other == 0 && (n_n == 0 || n_n <= __ldg(P.nmax + min(len, (uint32_t)kRegionBases)));
		const bool mate_ok = __shfl_xor_sync(0xFFFFFFFFu, (int)ok, 1); // (not inside '&&': every lane must shuffle)
		const bool pair_ok = ok && mate_ok;
		const uint32_t total = (len >= P.k) ? len - P.k + 1 : 0;
		uint32_t c_read = 0;      // the read's contig end (bestContig's return value)
		bool need_slow = false;   // read must go through the warp-wide path
		bool have_seed = false;
		uint64_t seed_pi = 0;
		uint32_t seed_p = 0;
		bool seed_fc = false;
		// a read whose windows are all classified keeps its tallies here until its lookups are back
		uint32_t r_ki = 0, r_found = 0, r_rec = 0, n_lookup = 0, r_cend = 0;
		bool r_same = false, classified = false;
		if (pair_ok && total) {
			need_slow = true;
			if (P.use_extension) {
				const uint32_t* Wr = G.W[lane];
				// seeds, two probes in flight
#pragma unroll 1
				for (uint32_t s0 = 0; s0 < (uint32_t)kGroupSeeds && !have_seed; s0 += 2) {
					// the two keys come out of ONE copy of the key code (rolled loop, results moved into named
					// registers); then both probes are issued back to back
					Key128 key0{0, 0}, key1{0, 0};
					uint64_t slot0 = 0, slot1 = 0;
					uint32_t sp0 = 0, sp1 = 0;
					bool fc0 = false, fc1 = false, act0 = false, act1 = false;
#pragma unroll 1
					for (uint32_t u = 0; u < 2; ++u) {
						const uint32_t sidx = s0 + u;
						const uint32_t sp = seed_window(sidx, total);
						const bool act = sidx < (uint32_t)kGroupSeeds && (clean || !window_invalid(G.s.m.INV[lane], sp, P.k));
						KeyHash kh{{0, 0}, 0, 0};
						if (act)
							kh = window_key_hash<KW>(Wr, sp, P.k, P.mask_hi, P.mask_lo);
						const uint64_t slot = hash_to_slot(kh.hash, P.nslots);
						if (u == 0) {
							key0 = kh.key, slot0 = slot, sp0 = sp, fc0 = kh.fwd_is_canonical != 0, act0 = act;
						} else {
							key1 = kh.key, slot1 = slot, sp1 = sp, fc1 = kh.fwd_is_canonical != 0, act1 = act;
						}
					}
					uint64_t hi0 = 0, lo0 = 0, pi0 = 0, hi1 = 0, lo1 = 0, pi1 = 0;
					uint32_t val0 = 0, val1 = 0;
					if (act0)
						load_slot(P.table, slot0, hi0, lo0, val0, pi0);
					if (act1)
						load_slot(P.table, slot1, hi1, lo1, val1, pi1);
					if (act0) {
						bool found = slot_matches<KW>(hi0, lo0, key0);
						if (!found && !slot_empty<KW>(hi0, lo0)) {
							const ChainResult cr = slot_chain_find<KW>(P.table, P.nslots, slot0, key0.hi, key0.lo);
							found = cr.found != 0;
							pi0 = cr.posinfo;
						}
						if (found) {
							have_seed = true;
							seed_pi = pi0;
							seed_p = sp0;
							seed_fc = fc0;
						}
					}
					if (act1 && !have_seed) {
						bool found = slot_matches<KW>(hi1, lo1, key1);
						if (!found && !slot_empty<KW>(hi1, lo1)) {
							const ChainResult cr = slot_chain_find<KW>(P.table, P.nslots, slot1, key1.hi, key1.lo);
							found = cr.found != 0;
							pi1 = cr.posinfo;
						}
						if (found) {
							have_seed = true;
							seed_pi = pi1;
							seed_p = sp1;
							seed_fc = fc1;
						}
					}
				}
				if (have_seed) {
					// diagonal of the seed.  "Text orientation": coordinate t of the read is text base gw0 + t
					// (t = read base t on the same strand, the complement of read base len-1-t otherwise);
					// window u in text orientation is read window u, resp. total-1-u.
					const bool c_fc = (seed_pi >> kPosBits) & 1ull;
					const uint64_t g = seed_pi & kPosMask;
					const uint32_t eidx = (uint32_t)(seed_pi >> 40);
					const bool same = seed_fc == c_fc;
					const uint32_t nw = (len + 15) >> 4;
					const uint32_t roff_s = (nw << 4) - len; // RC-stream coordinate of read base 0
					const uint32_t q0 = same ? 0u : roff_s;
					const int64_t D = same ? (int64_t)g - (int64_t)seed_p : (int64_t)g - (int64_t)((len - P.k - seed_p) + roff_s);
					const int64_t e_lo = (int64_t)__ldg(P.ct_end_g0 + eidx);
					const int64_t e_hi = e_lo + (int64_t)__ldg(P.ct_end_len + eidx);
					const int64_t gw0 = D + q0;
					// windows [u_lo, u_hi) lie inside the seed's contig end; the others (a read that runs over
					// the end of a contig end) are looked up.  Without the second-candidate logic only reads
					// inside one end are finished here.
					const bool all_inside = gw0 >= e_lo && gw0 + (int64_t)len <= e_hi;
					const bool in_text = gw0 >= 0 && gw0 + (int64_t)len <= (int64_t)P.ct_n_bases;
					if (in_text && (all_inside || P.lane_general)) {
						const uint32_t u_lo = e_lo > gw0 ? (uint32_t)min((int64_t)total, e_lo - gw0) : 0u;
						const int64_t hi_s = e_hi - (int64_t)P.k - gw0 + 1;
						const uint32_t u_hi = hi_s <= 0 ? 0u : (uint32_t)min((int64_t)total, hi_s);
						// the read against the contig text, 16 bases at a time: one mismatch bit per base
						uint32_t* Bm = G.s.m.BM[lane];
						uint32_t* Pm = G.PM[lane];
						const uint32_t nmw = (len + 31) >> 5;
						for (uint32_t i = 0; i <= nmw; ++i)
							Bm[i] = 0u;
						uint32_t any_mm = 0;
#pragma unroll kMismatchUnroll
						for (uint32_t j = 0; j < nw; ++j) {
							const uint32_t sw = same ? Wr[j] : rev2(~Wr[nw - 1 - j]);
							uint32_t m16 = mismatch16(sw, P.ct_T, D + 16 * (int64_t)j, (int64_t)P.ct_n_bases);
							const uint32_t b0 = 16 * j; // keep only stream coordinates in [q0, q0 + len)
							const uint32_t lo_b = q0 > b0 ? min(q0 - b0, 16u) : 0u;
							const uint32_t hi_b = q0 + len > b0 ? min(q0 + len - b0, 16u) : 0u;
							m16 &= (hi_b > lo_b) ? (((1u << hi_b) - 1u) & ~((1u << lo_b) - 1u)) : 0u;
							reinterpret_cast<uint16_t*>(Bm)[j] = (uint16_t)m16;
							any_mm |= m16;
						}
						const bool any_bad = any_mm != 0 || !clean;
						if (!any_bad || P.lane_general) {
							if (any_bad) {
								if (q0) // stream -> text orientation
									for (uint32_t i = 0; i < nmw; ++i)
										Bm[i] = __funnelshift_r(Bm[i], Bm[i + 1], q0);
								if (!clean) {
									const uint32_t* Ir = G.s.m.INV[lane];
									for (uint32_t i = 0; i < nmw; ++i) {
										const uint32_t v = same ? mask_bits_at(Ir, (int)(32 * i), (int)len)
										                        : __brev(mask_bits_at(Ir, (int)len - 32 - (int)(32 * i), (int)len));
										Pm[i] = v;
										Bm[i] |= v;
									}
									Pm[nmw] = 0u;
									dilate_row_call(Pm, nmw, P.k);
								}
								dilate_row_call(Bm, nmw, P.k);
							}
							// windows in text orientation, 32 at a time
							const uint64_t w0 = (uint64_t)gw0 >> 5;
							const uint32_t sh = (uint32_t)gw0 & 31u;
							const uint32_t nww = (total + 31) >> 5;
#pragma unroll 1
							for (uint32_t w = 0; w < nww; ++w) {
								const uint32_t left = total - 32u * w;
								const uint32_t range = left >= 32 ? 0xFFFFFFFFu : ((1u << left) - 1u);
								const uint32_t invw = clean ? 0u : (Pm[w] & range);
								const uint32_t badw = any_bad ? (Bm[w] & range) : 0u;
								const uint32_t ins = __funnelshift_r(__ldg(P.ct_TINS + w0 + w), __ldg(P.ct_TINS + w0 + w + 1), sh);
								const uint32_t uq = __funnelshift_r(__ldg(P.ct_TUNIQ + w0 + w), __ldg(P.ct_TUNIQ + w0 + w + 1), sh);
								const uint32_t inside = all_inside ? range : word_range_mask(u_lo, u_hi, w);
								const uint32_t res = inside & ~badw & ins; // all k bases equal an inserted window of this contig end
								const uint32_t lookup = range & ~invw & ~res;
								r_ki += __popc(invw);
								r_found += __popc(res);
								r_rec += __popc(res & uq);
								n_lookup += count_chunks(lookup);
								Pm[w] = lookup;
							}
							if (n_lookup == 0 || P.lane_general) {
								classified = true;
								need_slow = false;
								r_same = same;
								r_cend = __ldg(P.ct_end_cr + eidx);
							}
						}
					}
				}
			}
		} else if (pair_ok) {
			// shorter than k: bestContig returns 0 after an empty loop (maxjaccard 0)
			if (0.0 > P.j_index)
				pass++;
			else
				fail++;
		}
		// ---- stage 2b: the lookups of all 32 reads, flat over the warp in chunks of up to kChunkWindows
		// consecutive windows (rolling key within a chunk)
		const uint32_t my_lookups = classified ? n_lookup : 0u; // chunks
		if (__ballot_sync(0xFFFFFFFFu, my_lookups != 0)) {
			uint32_t incl_l = my_lookups;
#pragma unroll 1
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl_l, o);
				if (lane >= (uint32_t)o)
					incl_l += t;
			}
			const uint32_t excl_l = incl_l - my_lookups;
			const uint32_t all_lookups = __shfl_sync(0xFFFFFFFFu, incl_l, 31);
			G.rinfo[lane] = total | ((uint32_t)r_same << 16);
			G.rcend[lane] = r_cend;
			G.rcend2[lane] = 0;
			G.pcnt[lane] = 0;
			G.pcnt2[lane] = 0;
			// where base b of a left-aligned forward key sits, for b = k - 1 (the base a roll brings in)
			const uint32_t kin = P.k - 1;
			const uint32_t in_shift = 62u - 2u * (kin & 31u);
#pragma unroll 1
			for (uint32_t base = 0; base < all_lookups; base += kGroupListCap) {
				__syncwarp();
				if (my_lookups && excl_l < base + kGroupListCap && incl_l > base) {
					const uint32_t* Pm = G.PM[lane];
					const uint32_t nww = (total + 31) >> 5;
					uint32_t idx = excl_l;
					for (uint32_t w = 0; w < nww; ++w) {
						uint32_t m = Pm[w];
						while (m) {
							const uint32_t s = __ffs(m) - 1;
							const uint32_t t = m >> s;
							const uint32_t run = t == 0xFFFFFFFFu ? 32u : (uint32_t)__ffs(~t) - 1u;
							const uint32_t c = min(run, kChunkWindows);
							m &= ~((c >= 32u ? 0xFFFFFFFFu : ((1u << c) - 1u)) << s);
							if (idx >= base && idx < base + kGroupListCap)
								G.s.l.list[idx - base] = (uint16_t)((lane << 11) | ((32u * w + s) << 3) | (c - 1u));
							idx++;
						}
					}
				}
				__syncwarp();
				const uint32_t n_here = min((uint32_t)kGroupListCap, all_lookups - base);
#pragma unroll 1
				for (uint32_t f0 = 0; f0 < n_here; f0 += 32) {
					// 32 chunks at a time through the filter; the positives (rare) are collected and consulted
					// in the table afterwards, all at once, so that their DRAM round trips overlap
					if (lane == 0)
						G.s.l.npos = 0;
					__syncwarp();
					const uint32_t f = f0 + lane;
					if (f < n_here) {
						const uint32_t e = G.s.l.list[f];
						const uint32_t r = e >> 11, u0 = (e >> 3) & 255u, n = (e & 7u) + 1u;
						const uint32_t info = G.rinfo[r];
						const uint32_t tot = info & 0xFFFFu;
						const uint32_t p_lo = (info >> 16) ? u0 : tot - u0 - n; // read windows [p_lo, p_lo + n)
						const uint32_t* Wr = G.W[r];
						Key128 fw = extract_window<KW>(Wr, p_lo, P.mask_hi, P.mask_lo);
						Key128 rc = revcomp_key<KW>(fw, P.k);
#pragma unroll 1
						for (uint32_t i0 = 0; i0 < n; i0 += kLookupBatch) {
							// kLookupBatch filter words in flight
							uint32_t flo[kLookupBatch], fhi[kLookupBatch], sel[kLookupBatch];
#pragma unroll
							for (int jj = 0; jj < kLookupBatch; ++jj) {
								flo[jj] = fhi[jj] = 0xFFFFFFFFu;
								sel[jj] = 0u;
								if (i0 + jj < n) {
									const bool f_less = (fw.hi < rc.hi) || (fw.hi == rc.hi && fw.lo < rc.lo);
									const bool equal = (fw.hi == rc.hi) && (fw.lo == rc.lo);
									const Key128 key = equal ? palindrome_key(fw, (int)P.k) : (f_less ? fw : rc);
									if (P.bloom) {
										const BloomProbe bp = bloom_probe(key_hash<KW>(key), P.bloom_words);
										sel[jj] = bp.sel;
										bloom_load(P.bloom, bp, flo[jj], fhi[jj]);
									}
									// roll to the next window: read base p + k comes in
									const uint32_t q = p_lo + i0 + jj + P.k;
									const uint64_t nb = (Wr[q >> 4] >> (30u - 2u * (q & 15u))) & 3u;
									if (KW == 1) {
										fw.hi = (fw.hi << 2) | (nb << in_shift);
										rc.hi = ((rc.hi >> 2) | ((3ull - nb) << 62)) & P.mask_hi;
									} else {
										fw.hi = (fw.hi << 2) | (fw.lo >> 62);
										fw.lo <<= 2;
										if (kin < 32)
											fw.hi |= nb << in_shift;
										else
											fw.lo |= nb << in_shift;
										rc.lo = ((rc.lo >> 2) | (rc.hi << 62)) & P.mask_lo;
										rc.hi = ((rc.hi >> 2) | ((3ull - nb) << 62)) & P.mask_hi;
									}
								}
							}
#pragma unroll
							for (int jj = 0; jj < kLookupBatch; ++jj) {
								if (i0 + jj < n && bloom_test(sel[jj], flo[jj], fhi[jj]))
									G.s.l.pos[atomicAdd(&G.s.l.npos, 1u)] = (uint16_t)((r << 8) | (p_lo + i0 + jj));
							}
						}
					}
					__syncwarp();
					const uint32_t n_pos = G.s.l.npos;
#pragma unroll 1
					for (uint32_t i = lane; i < n_pos; i += 32) {
						const uint32_t e = G.s.l.pos[i];
						const uint32_t r = e >> 8;
						const KeyHash kh = window_key_hash<KW>(G.W[r], e & 255u, P.k, P.mask_hi, P.mask_lo);
						const uint64_t slot = hash_to_slot(kh.hash, P.nslots);
						uint64_t hi, lo, pi;
						uint32_t val;
						load_slot(P.table, slot, hi, lo, val, pi);
						bool found = slot_matches<KW>(hi, lo, kh.key);
						if (!found && !slot_empty<KW>(hi, lo)) {
							const ChainResult cr = slot_chain_find<KW>(P.table, P.nslots, slot, kh.key.hi, kh.key.lo);
							found = cr.found != 0;
							val = cr.val;
						}
						if (found) {
							atomicAdd(&G.pcnt[r], val ? 0x10001u : 1u);
							if (val && val != G.rcend[r]) {
								// a vote for another contig end: one more candidate is tracked
								const uint32_t old = atomicCAS(&G.rcend2[r], 0u, val);
								if (old == 0u || old == val)
									atomicAdd(&G.pcnt2[r], 1u);
								else
									atomicOr(&G.pcnt2[r], 0x80000000u);
							}
						}
					}
					__syncwarp();
				}
			}
			__syncwarp();
		}
		if (classified) {
			// bestContig's tail (Arcs.cpp:996-1012): the recorded windows voted for the seed's contig end,
			// except the looked-up ones that voted for a second one
			uint32_t v1 = r_rec, v2 = 0, c2 = 0;
			bool third = false;
			if (my_lookups) {
				const uint32_t c = G.pcnt[lane], d = G.pcnt2[lane];
				third = (d >> 31) != 0 || (d != 0 && !P.lane_general);
				v2 = d & 0x7FFFFFFFu;
				c2 = G.rcend2[lane];
				r_found += c & 0xFFFFu;
				r_rec += c >> 16;
				v1 = r_rec - v2;
			}
			if (third) {
				// votes for three contig ends: the general path takes the read from scratch
				need_slow = true;
			} else {
				st.kv += total - r_ki;
				st.ki += r_ki;
				st.found += r_found;
				st.rec += r_rec;
				st.dups += r_found - r_rec;
				// argmax, ties -> smallest contig end (std::map order + strict '<', Arcs.cpp:998-1003)
				const uint32_t best = max(v1, v2);
				const uint32_t best_c = (v2 > v1 || (v2 == v1 && c2 < r_cend)) ? c2 : r_cend;
				bool passed;
				if (best == 0)
					passed = 0.0 > P.j_index;
				else
					passed = best >= __ldg(P.jmin + total);
				if (passed) {
					pass++;
					c_read = best ? best_c : 0u;
				} else {
					fail++;
				}
			}
		}
		// ---- stage 3: the pair rule for pairs whose mates are both finished (one lane per pair,
		// Arcs.cpp:1280); every other valid pair goes to the work list of map_slow_kernel
		const uint32_t c_mate = __shfl_xor_sync(0xFFFFFFFFu, c_read, 1);
		const bool mate_slow = __shfl_xor_sync(0xFFFFFFFFu, (int)need_slow, 1);
		const bool even = (lane & 1) == 0;
		const bool defer = exists && even && (pair_long || (pair_ok && (need_slow || mate_slow)));
		const uint32_t defer_mask = __ballot_sync(0xFFFFFFFFu, defer);
		if (defer_mask) {
			uint32_t base = 0;
			if (lane == 0)
				base = atomicAdd(P.work_count, (uint32_t)__popc(defer_mask));
			base = __shfl_sync(0xFFFFFFFFu, base, 0);
			// the record is written by the pair's even lane; it needs the odd lane's read too
			const uint32_t my_meta = (have_seed ? 1u : 0u) | (seed_fc ? 2u : 0u) | (seed_p << 8);
			const uint32_t m_meta = __shfl_xor_sync(0xFFFFFFFFu, my_meta, 1);
			const uint32_t m_lo = __shfl_xor_sync(0xFFFFFFFFu, (uint32_t)seed_pi, 1);
			const uint32_t m_hi = __shfl_xor_sync(0xFFFFFFFFu, (uint32_t)(seed_pi >> 32), 1);
			const uint32_t m_off = __shfl_xor_sync(0xFFFFFFFFu, off, 1);
			const uint32_t m_len = __shfl_xor_sync(0xFFFFFFFFu, len, 1);
			if (defer) {
				uint4* dst = reinterpret_cast<uint4*>(P.work + base + __popc(defer_mask & ((1u << lane) - 1u)));
				const uint32_t s0 = pair_long ? kMateUnknown : (need_slow ? kMateSlow : c_read);
				const uint32_t s1 = pair_long ? kMateUnknown : (mate_slow ? kMateSlow : c_mate);
				dst[0] = make_uint4(my_pair, off, m_off, m_off + m_len);
				dst[1] = make_uint4(s0, s1, my_meta, m_meta);
				dst[2] = make_uint4((uint32_t)seed_pi, m_lo, (uint32_t)(seed_pi >> 32), m_hi);
			}
		}
		if (exists && even && !defer) {
			uint32_t out = 0;
			if (!pair_ok)
				invalid++;
			if (c_read != 0 && c_read == c_mate) {
				stored++;
				out = c_read;
				const uint32_t cc = (P.remap && out < P.n_remap) ? P.remap[out] : out;
				imap_add_call(P.imap, P.imap_mask, P.imap_count, P.barcode_id[my_pair], (cc - 1) >> 1, (cc & 1u), (cc & 1u) ^ 1u);
			} else {
				nogood++;
			}
			if (P.conreci_out)
				P.conreci_out[my_pair] = (int32_t)out;
		}
	}
	flush_counters(P, lane, st, pass, fail, stored, invalid, nogood, false);
}

// ---- the pairs the group kernel deferred: one warp per pair, general path -------------------
template <int KW>
__global__ void __launch_bounds__(kMapThreads, kMapMinBlocks) map_slow_kernel(MapParams P)
{
	__shared__ WarpRegion regions[kMapWarps][2];
	__shared__ uint32_t Ms[kMapWarps][kMWords];
	__shared__ uint16_t lists[kMapWarps][kRegionBases];
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warp = threadIdx.x >> 5;
	LaneStats st{0, 0, 0, 0, 0};
	PairCounters pc{0, 0, 0, 0, 0, false};
	const uint32_t n_work = *P.work_count;
	const uint32_t nwarps = gridDim.x * kMapWarps;
	__shared__ WorkRecord recs[kMapWarps];
	// work items are handed out by a ticket counter (they differ a lot in cost); a warp holds its next two tickets
	// so that the software pipeline below knows what comes: the record two items ahead is prefetched by address; the
	// next item's record (an L1 hit by now) tells which lines ITS mates, contig text and masks live in, and those
	// are prefetched while the current item is processed.
	auto take = [&]() -> uint32_t {
#ifndef ARKS_STATIC_GROUPS
		uint32_t t = 0;
		if (lane == 0)
			t = atomicAdd(P.work_count + 2, 1u);
		return nwarps + __shfl_sync(0xFFFFFFFFu, t, 0);
#else
		return 0xFFFFFFFFu;
#endif
	};
	uint32_t i = blockIdx.x * kMapWarps + warp;
#ifndef ARKS_STATIC_GROUPS
	uint32_t i1 = i < n_work ? take() : 0xFFFFFFFFu, i2 = i1 < n_work ? take() : 0xFFFFFFFFu;
#else
	uint32_t i1 = i + nwarps, i2 = i + 2 * nwarps;
#endif
#pragma unroll 1
	for (; i < n_work;) {
		if (i2 < n_work && lane == 0)
			prefetch_l1(P.work + i2);
		if (i1 < n_work)
			prefetch_item(P.work[i1], P, lane);
		__syncwarp();
		if (lane < 4)
			reinterpret_cast<uint4*>(&recs[warp])[lane] = reinterpret_cast<const uint4*>(P.work + i)[lane];
		__syncwarp();
		warp_process_pair<KW>(recs[warp], P, regions[warp], Ms[warp], lists[warp], lane, st, pc);
		i = i1;
		i1 = i2;
#ifndef ARKS_STATIC_GROUPS
		i2 = i2 < n_work ? take() : 0xFFFFFFFFu;
#else
		i2 = i2 + nwarps;
#endif
	}
	const bool l0 = lane == 0;
	flush_counters(P, lane, st, l0 ? pc.pass : 0, l0 ? pc.fail : 0, l0 ? pc.stored : 0, l0 ? pc.invalid : 0, l0 ? pc.nogood : 0,
	    pc.overflow);
}

} // namespace arks
