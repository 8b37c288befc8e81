// arks_device.cuh -- device-side building blocks shared by the index-build and the
// read-lookup kernels (sm_100a).
//
//  * ASCII -> 2-bit packing of a base region into shared memory (forward stream W,
//    reverse-complement stream RC, invalid-base bitmask INV), 16 bases per thread.
//  * O(1) extraction of the canonical key of any window of the region from the two
//    packed streams (no rolling state, every lane works on an independent window).
//  * The reference's key semantics: ReadsProcessor::prepSeq, Common/ReadsProcessor.cpp:
//    376-535 -- canonical = lexicographic min(window, revcomp) in MSB-first 2-bit
//    packing, no key if any base is not ACGTacgt, and the deterministic garbage key for
//    windows equal to their own reverse complement (:503-534).
//  * Slot hash + table probe.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace arks {

constexpr uint64_t kEmptyKey = ~0ull;          // all-ones 128-bit key = empty slot
constexpr uint64_t kEmptyW = ~0ull;            // build-time bookkeeping word, unset
constexpr uint32_t kMultiFlag = 0x80000000u;   // build-time: key seen in >1 contig end
constexpr int kMaxTrack = 32;                  // distinct contig ends tracked per read (one per lane)

// ---------------------------------------------------------------------------------
// Region packing.  A "region" is `len` consecutive bases.  Thread g of the cooperating
// group packs bases [16g, 16g+16) into W[g] (base 16g in bits 31-30).  Positions >= len
// pack as code 0 and are not flagged invalid.  INV bit i (LSB-first, word i>>5) is set
// iff base i is not one of ACGTacgt.
// ---------------------------------------------------------------------------------

// Loads the 16 bytes [a, a+16) as four little-endian words using aligned 32-bit loads
// that never touch a word lying entirely outside [lo, hi).
__device__ __forceinline__ void load16_unaligned(const char* a, const char* hi, uint32_t x[4])
{
	uintptr_t ua = reinterpret_cast<uintptr_t>(a);
	const uint32_t* a4 = reinterpret_cast<const uint32_t*>(ua & ~uintptr_t(3));
	const uint32_t sh = (uint32_t)(ua & 3) * 8;
	uint32_t w[5];
#pragma unroll
	for (int j = 0; j < 5; ++j)
		// (streaming load: every base is read exactly once, so it should not push the filter and the contig text out
		// of L2 -- +1.8 % on configs[1] against a plain load)
		w[j] = (reinterpret_cast<const char*>(a4 + j) < hi) ? __ldcs(a4 + j) : 0u;
#pragma unroll
	for (int j = 0; j < 4; ++j)
		x[j] = __funnelshift_r(w[j], w[j + 1], sh);
}

// 4 ASCII bases (little-endian word, first base in the low byte) ->
//   returns 8 bits MSB-first packed (first base in bits 7-6),
//   *bad4 gets a 4-bit mask (bit i = base i is not ACGTacgt).
__device__ __forceinline__ uint32_t pack4(uint32_t x, uint32_t nvalid, uint32_t* bad4)
{
	// A=0x41 C=0x43 G=0x47 T=0x54 (and lower case): (c>>1)&3 -> A0 C1 G3 T2; xor bit 2 fixes G/T.
	uint32_t code = ((x >> 1) & 0x03030303u) ^ ((x >> 2) & 0x01010101u);
	// expected lower-case letter for each code through PRMT as a 4-entry byte LUT
	uint32_t t = code | (code >> 4);
	uint32_t sel = __byte_perm(t, 0, 0x4420);          // nibbles c0,c1,c2,c3
	uint32_t expect = __byte_perm(0x74676361u, 0, sel); // 'a','c','g','t'
	uint32_t diff = (x | 0x20202020u) ^ expect;
	// bytes at positions >= nvalid are padding: force them to valid code 0
	uint32_t keep = nvalid >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nvalid)) - 1u);
	diff &= keep;
	code &= keep;
	uint32_t b = 0;
	if (diff) {
		uint32_t nz = (((diff & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | diff) & 0x80808080u;
		b = ((nz >> 7) & 1u) | ((nz >> 14) & 2u) | ((nz >> 21) & 4u) | ((nz >> 28) & 8u);
	}
	*bad4 = b;
	return (code * 0x40100401u) >> 24;
}

// counts bytes equal to 'N'/'n' among the first nvalid bytes of x
__device__ __forceinline__ uint32_t count_n4(uint32_t x, uint32_t nvalid)
{
	uint32_t d = (x | 0x20202020u) ^ 0x6E6E6E6Eu;
	uint32_t nz = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u; // 0x80 where byte != 'n'
	uint32_t isn = (~nz) & 0x80808080u;
	uint32_t keep = nvalid >= 4 ? 0x80808080u : (((1u << (8 * nvalid)) - 1u) & 0x80808080u);
	return __popc(isn & keep);
}

// Packs group g (bases [16g, 16g+16) of the region that starts at src and has len bases).
// Returns the packed word; inv16 = invalid mask of the 16 bases; n_n = number of N/n;
// n_other = number of invalid bases that are not N/n.
__device__ __forceinline__ uint32_t
pack_group(const char* src, uint32_t len, uint32_t g, uint32_t* inv16, uint32_t* n_n, uint32_t* n_other)
{
	uint32_t x[4];
	uint32_t base = 16 * g;
	load16_unaligned(src + base, src + len, x);
	uint32_t remain = len - base; // > 0 guaranteed by caller
	uint32_t word = 0, inv = 0, nn = 0;
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		uint32_t nv = remain > 4u * j ? remain - 4u * j : 0u;
		uint32_t bad;
		uint32_t p = pack4(x[j], nv, &bad);
		word |= p << (24 - 8 * j);
		inv |= bad << (4 * j);
		if (bad)
			nn += count_n4(x[j], nv);
	}
	*inv16 = inv;
	*n_n = nn;
	*n_other = __popc(inv) - nn;
	return word;
}

// same as pack_group with the four 4-base steps rolled into a loop (a quarter of the code: the
// lane-per-read lookup kernel is instruction-cache bound)
__device__ __forceinline__ uint32_t
pack_group_compact(const char* src, uint32_t len, uint32_t g, uint32_t* inv16, uint32_t* n_n, uint32_t* n_other)
{
	uint32_t x[4];
	const uint32_t base = 16 * g;
	load16_unaligned(src + base, src + len, x);
	const uint32_t remain = len - base; // > 0 guaranteed by caller
	uint32_t word = 0, inv = 0, nn = 0;
#pragma unroll 1
	for (uint32_t j = 0; j < 4; ++j) {
		// x[j] through selects so that the array stays in registers
		const uint32_t xj = j == 0 ? x[0] : (j == 1 ? x[1] : (j == 2 ? x[2] : x[3]));
		const uint32_t nv = remain > 4u * j ? remain - 4u * j : 0u;
		uint32_t bad;
		const uint32_t p = pack4(xj, nv, &bad);
		word = (word << 8) | p;
		inv |= bad << (4 * j);
		if (bad)
			nn += count_n4(xj, nv);
	}
	*inv16 = inv;
	*n_n = nn;
	*n_other = __popc(inv) - nn;
	return word;
}

// reverse the order of the sixteen 2-bit groups of x
__device__ __forceinline__ uint32_t rev2(uint32_t x)
{
	uint32_t y = __brev(x);
	return ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
}

// ---------------------------------------------------------------------------------
// Window extraction.  S is a packed stream (MSB-first, 16 bases per word) with at
// least 5 readable words after the word that holds base p.
// ---------------------------------------------------------------------------------
struct Key128
{
	uint64_t hi, lo; // big-endian: hi holds bases 0..31 (base 0 in bits 63-62)
};

template <int KW>
__device__ __forceinline__ Key128 extract_window(const uint32_t* S, uint32_t p, uint64_t mask_hi, uint64_t mask_lo)
{
	uint32_t wi = p >> 4, s = (p & 15u) * 2u;
	uint32_t w0 = S[wi], w1 = S[wi + 1], w2 = S[wi + 2];
	uint32_t x0 = __funnelshift_l(w1, w0, s);
	uint32_t x1 = __funnelshift_l(w2, w1, s);
	Key128 r;
	r.hi = (((uint64_t)x0 << 32) | x1) & mask_hi;
	r.lo = 0;
	if (KW == 2) {
		uint32_t w3 = S[wi + 3], w4 = S[wi + 4];
		uint32_t x2 = __funnelshift_l(w3, w2, s);
		uint32_t x3 = __funnelshift_l(w4, w3, s);
		r.lo = (((uint64_t)x2 << 32) | x3) & mask_lo;
	}
	return r;
}

// true iff any of the k bases of window p is invalid (INV is LSB-first, 2 readable words
// after the word holding bit p)
__device__ __forceinline__ bool window_invalid(const uint32_t* INV, uint32_t p, uint32_t k)
{
	uint32_t wi = p >> 5, s = p & 31u;
	uint32_t w0 = INV[wi], w1 = INV[wi + 1], w2 = INV[wi + 2];
	uint32_t lo = __funnelshift_r(w0, w1, s);
	uint32_t hi = __funnelshift_r(w1, w2, s);
	uint64_t bits = ((uint64_t)hi << 32) | lo;
	uint64_t m = k >= 64 ? ~0ull : ((1ull << k) - 1ull);
	return (bits & m) != 0ull;
}

// base i (0..k-1) of a left-aligned packed k-mer
__device__ __forceinline__ uint32_t key_base(const Key128& f, int i)
{
	return i < 32 ? (uint32_t)(f.hi >> (62 - 2 * i)) & 3u : (uint32_t)(f.lo >> (62 - 2 * (i - 32))) & 3u;
}

__device__ __forceinline__ void key_set_byte(Key128& r, int o, uint32_t byte)
{
	if (o < 8)
		r.hi |= (uint64_t)byte << (56 - 8 * o);
	else
		r.lo |= (uint64_t)byte << (56 - 8 * (o - 8));
}

// The key the reference returns for a window that equals its own reverse complement
// (Common/ReadsProcessor.cpp:503-534): bytes [0,h) forward, byte h left at 0, later full
// bytes packed from a cursor that advances 3 bases per byte, hanging byte = one base in
// bits 7-6.  f = forward packing.  Rare path.
__device__ __noinline__ Key128 palindrome_key(Key128 f, int k)
{
	const int nb = (k + 3) >> 2;
	const int h = (k >> 3) + ((k & 7) != 0);
	const int hang = k & 3;
	Key128 r{0, 0};
	for (int o = 0; o < h; ++o) {
		uint32_t byte = 0;
		for (int j = 0; j < 4; ++j)
			byte = (byte << 2) | key_base(f, 4 * o + j);
		key_set_byte(r, o, byte);
	}
	int idx = 4 * h, o = h + 1;
	while (o + (hang ? 1 : 0) < nb) {
		uint32_t byte = (key_base(f, idx) << 6) | (key_base(f, idx + 1) << 4) | (key_base(f, idx + 2) << 2) |
		                key_base(f, idx + 3);
		key_set_byte(r, o, byte);
		idx += 3;
		o += 1;
	}
	if (hang && o < nb) {
		uint32_t c = (idx < k - 1) ? key_base(f, idx + 1) : key_base(f, k - 1);
		key_set_byte(r, o, c << 6);
	}
	return r;
}

// canonical key of window p of a region of L bases whose forward stream is W and whose
// reverse-complement stream is RC (RC base i = complement of region base Lp-1-i where
// Lp = 16*nwords is the padded length, so window p of the region is window
// (Lp - k - p) of RC).
template <int KW>
__device__ __forceinline__ Key128
canonical_key(const uint32_t* W, const uint32_t* RC, uint32_t p, uint32_t k, uint32_t Lp, uint64_t mask_hi, uint64_t mask_lo,
    bool* fwd_is_canonical = nullptr)
{
	Key128 f = extract_window<KW>(W, p, mask_hi, mask_lo);
	Key128 r = extract_window<KW>(RC, Lp - k - p, mask_hi, mask_lo);
	bool f_less = (f.hi < r.hi) || (f.hi == r.hi && f.lo < r.lo);
	bool equal = (f.hi == r.hi) && (f.lo == r.lo);
	if (fwd_is_canonical)
		*fwd_is_canonical = f_less || equal;
	if (equal)
		return palindrome_key(f, (int)k);
	return f_less ? f : r;
}

// ---------------------------------------------------------------------------------
// Slot hash.  Not result-bearing (the full key is stored and compared).
// ---------------------------------------------------------------------------------
template <int KW>
__device__ __forceinline__ uint64_t key_hash(const Key128& key)
{
	uint64_t h = key.hi * 0x9E3779B97F4A7C15ull;
	if (KW == 2)
		h ^= (key.lo + 0x7F4A7C159E3779B9ull) * 0xC2B2AE3D27D4EB4Full;
	h ^= h >> 29;
	h *= 0xBF58476D1CE4E5B9ull;
	return h;
}

__device__ __forceinline__ uint64_t hash_to_slot(uint64_t h, uint64_t nslots)
{
	return __umul64hi(h, nslots);
}

// Table slot, 32 B (one DRAM sector) for every k:
//   [0]  key.hi          [8]  key.lo (unused, left all-ones, when k <= 32)
//   [16] w               build time: bookkeeping word; frozen: low 32 bits = value
//                        (contig-end record, 0 = seen in several ends), high 32 bits = 0
//   [24] posinfo         where the claiming occurrence sits in the packed contig text:
//                        bits 0-38 global base coordinate of the window, bit 39 = canonical
//                        key is the forward strand of the contig, bits 40-63 = global index
//                        of the contig end
constexpr int kSlotBytes = 32;
constexpr uint32_t kMiss = 0xFFFFFFFFu;
constexpr int kPosBits = 39;
constexpr uint64_t kPosMask = (1ull << kPosBits) - 1ull;

__device__ __forceinline__ void
load_slot(const uint8_t* table, uint64_t slot, uint64_t& hi, uint64_t& lo, uint32_t& val, uint64_t& posinfo)
{
	const uint8_t* p = table + slot * kSlotBytes;
	uint64_t a, b, c, d;
	asm("ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
	hi = a;
	lo = b;
	val = (uint32_t)c;
	posinfo = d;
}

// Membership prefilter in front of the table: one 64-bit word per key, two bits in each 32-bit half
// (register-blocked Bloom filter).  Sized to stay resident in the 126 MB L2 for drafts up to a few
// hundred Mbp, so that the lookup of a read k-mer that is NOT in the draft (every window that
// overlaps a sequencing error) costs an L2 hit instead of a DRAM line.  No false negatives, so the
// table is consulted only on a positive and results do not depend on the filter.
struct BloomProbe
{
	uint64_t word;
	uint32_t sel; // four 5-bit bit indices: two in the low half of the word, two in the high half
};

__device__ __forceinline__ BloomProbe bloom_probe(uint64_t key_hash_value, uint64_t n_words)
{
	const uint64_t h2 = (key_hash_value ^ (key_hash_value >> 32)) * 0x9E3779B97F4A7C15ull;
	BloomProbe b;
	b.word = __umul64hi(h2, n_words);
	b.sel = (uint32_t)h2 >> 12;
	return b;
}

__device__ __forceinline__ uint32_t bloom_mask_lo(uint32_t sel)
{
	return (1u << (sel & 31u)) | (1u << ((sel >> 5) & 31u));
}

__device__ __forceinline__ uint32_t bloom_mask_hi(uint32_t sel)
{
	return (1u << ((sel >> 10) & 31u)) | (1u << ((sel >> 15) & 31u));
}

__device__ __forceinline__ bool bloom_test(uint32_t sel, uint32_t lo, uint32_t hi)
{
	const uint32_t m_lo = bloom_mask_lo(sel), m_hi = bloom_mask_hi(sel);
	return (lo & m_lo) == m_lo && (hi & m_hi) == m_hi;
}

// L2 residency policies (createpolicy folds into a constant descriptor)
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
	uint64_t pol;
	asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}

__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
	uint64_t pol;
	asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}

__device__ __forceinline__ void bloom_load(const unsigned long long* bloom, const BloomProbe& b, uint32_t& lo, uint32_t& hi)
{
	asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;"
	    : "=r"(lo), "=r"(hi)
	    : "l"(bloom + b.word), "l"(l2_policy_evict_last()));
}

__device__ __forceinline__ bool bloom_maybe(const unsigned long long* bloom, const BloomProbe& b)
{
	uint32_t lo, hi;
	bloom_load(bloom, b, lo, hi);
	return bloom_test(b.sel, lo, hi);
}

template <int KW>
__device__ __forceinline__ bool slot_matches(uint64_t hi, uint64_t lo, const Key128& key)
{
	return hi == key.hi && (KW == 1 || lo == key.lo);
}

template <int KW>
__device__ __forceinline__ bool slot_empty(uint64_t hi, uint64_t lo)
{
	return hi == kEmptyKey && (KW == 1 || lo == kEmptyKey);
}

// reverse the 32 two-bit groups of a 64-bit word
__device__ __forceinline__ uint64_t rev2_64(uint64_t x)
{
	return ((uint64_t)rev2((uint32_t)x) << 32) | rev2((uint32_t)(x >> 32));
}

// reverse complement of a left-aligned packed k-mer (same result as reading the
// reverse-complement stream)
template <int KW>
__device__ __forceinline__ Key128 revcomp_key(const Key128& f, uint32_t k)
{
	Key128 r;
	if (KW == 1) {
		uint64_t a = rev2_64(~f.hi); // pad bases (complemented to T) lead: shift them out
		r.hi = a << (2 * (32 - k));
		r.lo = 0;
	} else {
		uint64_t a = rev2_64(~f.lo), b = rev2_64(~f.hi); // (a:b) = revcomp of the 64-base padded string
		uint32_t s = 2 * (64 - k);                        // 0 .. 62
		r.hi = s ? (a << s) | (b >> (64 - s)) : a;
		r.lo = b << s;
	}
	return r;
}

// canonical key from the forward packing alone
template <int KW>
__device__ __forceinline__ Key128 canonical_from_forward(const Key128& f, uint32_t k, bool* fwd_is_canonical)
{
	Key128 r = revcomp_key<KW>(f, k);
	bool f_less = (f.hi < r.hi) || (f.hi == r.hi && f.lo < r.lo);
	bool equal = (f.hi == r.hi) && (f.lo == r.lo);
	*fwd_is_canonical = f_less || equal;
	if (equal)
		return palindrome_key(f, (int)k);
	return f_less ? f : r;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v)
{
	return __reduce_add_sync(0xFFFFFFFFu, v);
}

} // namespace arks
