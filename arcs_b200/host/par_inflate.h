// par_inflate.h -- multi-threaded decompression of ONE gzip member (the usual .fq.gz).
//
// A DEFLATE stream is sequential twice over: blocks start at arbitrary bit positions, and every block
// may copy from the 32 KB of output before it.  Both are worked around the way pugz does it
// (Kerbiriou & Chikhi, 2019), restated here from the idea:
//
//  1. the compressed bytes are cut into chunks; for the start of each chunk a block boundary is FOUND by
//     trying bit positions: a dynamic-Huffman block header whose code-length, literal/length and distance
//     codes are all complete, followed by two blocks that decode to plausible text (FASTQ bytes) -- a few
//     thousand candidates, each rejected after a handful of bits;
//  2. every chunk is decoded on its own thread into 16-bit symbols: a byte, or "the byte that was at
//     position w of the unknown 32 KB window in front of this chunk" (copies propagate those);
//  3. the windows are resolved chunk after chunk (32 K table look-ups each), then every chunk is translated to
//     bytes in parallel, CRC-32s are combined, and the member's CRC-32 / ISIZE are checked at the end.
//
// A chunk has to END exactly where the next one was found to start; if it does not (a false boundary), if
// no boundary is found, or if the file is anything but one plain member, the rest is decoded by the
// sequential decoder of fast_inflate.h from the last position that is known to be good.  The result is
// therefore always what a sequential decoder gives; tests/test_host_cpu.py checks that against zlib.
#pragma once
#include "fast_inflate.h"

#include <atomic>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif
#include <cstdlib>
#include <functional>
#include <memory>
#include <new>
#include <thread>

namespace arks_host {

namespace pinf {

// ---- bits at an absolute position of a memory range (LSB first; reads past the end see zeros) ----
struct Bits
{
	const uint8_t* base;
	size_t nbytes;
	uint64_t pos; // bit position
	uint64_t peek() const
	{
		const size_t b = (size_t)(pos >> 3);
		uint64_t w = 0;
		if (b + 8 <= nbytes)
			memcpy(&w, base + b, 8);
		else if (b < nbytes)
			memcpy(&w, base + b, nbytes - b);
		return w >> (pos & 7);
	}
	uint32_t get(int n)
	{
		const uint32_t v = (uint32_t)(peek() & ((1ull << n) - 1));
		pos += (uint64_t)n;
		return v;
	}
	bool past_end() const { return pos > (uint64_t)nbytes * 8; }
};

// ---- Huffman tables of one block (single level: 15-bit direct look-up would be 32 K entries; two levels as in
// fast_inflate.h keep it in L1) ----
struct Tables
{
	static constexpr int kLit = 11, kDist = 8;
	std::vector<uint32_t> lit, dist;
	// complete = every code is used (what compressors emit); required when hunting for block starts
	static bool build(const uint8_t* lens, int n, int primary, std::vector<uint32_t>& tab, bool need_complete, bool allow_single)
	{
		int count[16] = { 0 };
		for (int i = 0; i < n; ++i)
			count[lens[i]]++;
		count[0] = 0;
		int left = 1, used = 0;
		for (int l = 1; l <= 15; ++l) {
			left = (left << 1) - count[l];
			used += count[l];
			if (left < 0)
				return false;
		}
		if (need_complete && left != 0 && !(allow_single && used <= 1))
			return false;
		uint32_t next[16], code = 0;
		for (int l = 1; l <= 15; ++l) {
			code = (code + count[l - 1]) << 1;
			next[l] = code;
		}
		tab.assign((size_t)1 << primary, 0);
		uint8_t sub_max[1 << 11] = { 0 };
		uint32_t codes[320];
		for (int i = 0; i < n; ++i) {
			const int l = lens[i];
			if (!l)
				continue;
			uint32_t c = next[l]++, r = 0;
			for (int b = 0; b < l; ++b) {
				r = (r << 1) | (c & 1);
				c >>= 1;
			}
			codes[i] = r;
			if (l > primary) {
				uint8_t& m = sub_max[r & ((1u << primary) - 1)];
				if (l > m)
					m = (uint8_t)l;
			}
		}
		for (uint32_t p = 0; p < (1u << primary); ++p)
			if (sub_max[p]) {
				const int sb = sub_max[p] - primary;
				tab[p] = ((uint32_t)tab.size() << 8) | 0x80u | (uint32_t)sb;
				tab.resize(tab.size() + ((size_t)1 << sb), 0);
			}
		for (int i = 0; i < n; ++i) {
			const int l = lens[i];
			if (!l)
				continue;
			if (l <= primary) {
				for (uint32_t k = codes[i]; k < (1u << primary); k += 1u << l)
					tab[k] = ((uint32_t)i << 8) | (uint32_t)l;
			} else {
				const uint32_t e = tab[codes[i] & ((1u << primary) - 1)];
				const int sb = (int)(e & 0x7f);
				const uint32_t base = e >> 8;
				for (uint32_t k = codes[i] >> primary; k < (1u << sb); k += 1u << (l - primary))
					tab[base + k] = ((uint32_t)i << 8) | (uint32_t)(l - primary);
			}
		}
		return true;
	}
	// reads a dynamic block's code descriptions at `in` (positioned behind the 3 header bits)
	bool read_dynamic(Bits& in, bool strict)
	{
		const int hlit = (int)in.get(5) + 257, hdist = (int)in.get(5) + 1, hclen = (int)in.get(4) + 4;
		if (hlit > 286 || hdist > 30)
			return false;
		static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
		uint8_t cl[19] = { 0 };
		for (int i = 0; i < hclen; ++i)
			cl[order[i]] = (uint8_t)in.get(3);
		std::vector<uint32_t> cltab;
		if (!build(cl, 19, 7, cltab, strict, false))
			return false;
		uint8_t lens[286 + 30] = { 0 };
		int i = 0;
		while (i < hlit + hdist) {
			const uint32_t e = cltab[in.peek() & 127];
			if (!e || (e & 0x80))
				return false;
			in.pos += e & 0x7f;
			const int sym = (int)(e >> 8);
			if (sym < 16) {
				lens[i++] = (uint8_t)sym;
				continue;
			}
			int rep, val = 0;
			if (sym == 16) {
				if (i == 0)
					return false;
				val = lens[i - 1];
				rep = 3 + (int)in.get(2);
			} else if (sym == 17) {
				rep = 3 + (int)in.get(3);
			} else {
				rep = 11 + (int)in.get(7);
			}
			if (i + rep > hlit + hdist)
				return false;
			while (rep--)
				lens[i++] = (uint8_t)val;
		}
		if (lens[256] == 0 || in.past_end())
			return false;
		return build(lens, hlit, kLit, lit, strict, false) && build(lens + hlit, hdist, kDist, dist, strict, true);
	}
	bool set_fixed()
	{
		uint8_t l[288], d[30];
		for (int i = 0; i < 288; ++i)
			l[i] = i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : 8));
		for (int i = 0; i < 30; ++i)
			d[i] = 5;
		return build(l, 288, kLit, lit, false, false) && build(d, 30, kDist, dist, false, false);
	}
};

// grow-only buffer of 16-bit symbols whose new part is NOT cleared (a std::vector would zero-fill it)
struct SymBuf
{
	uint16_t* p = nullptr;
	size_t cap = 0;
	SymBuf() = default;
	SymBuf(const SymBuf&) = delete;
	SymBuf& operator=(const SymBuf&) = delete;
	SymBuf(SymBuf&& o) noexcept
	  : p(o.p)
	  , cap(o.cap)
	{
		o.p = nullptr;
		o.cap = 0;
	}
	~SymBuf() { free(p); }
	void ensure(size_t n)
	{
		if (n <= cap)
			return;
		void* q = realloc(p, n * sizeof(uint16_t));
		if (!q)
			throw std::bad_alloc();
		p = (uint16_t*)q;
		cap = n;
	}
};

inline bool plausible_text(uint32_t c)
{
	return c == '\n' || c == '\t' || c == '\r' || (c >= 32 && c < 127);
}

// Decodes one block at in.pos into 16-bit symbols appended to `out` (values >= 256: position in the window
// in front of out[0], see the header).  `first_is_known_start`: out[0] is the first byte of the member, so a
// copy from before it is an error.  `text_only`: every literal must be plausible text (block-start hunting).
// Returns false on anything invalid; *final tells whether it was the last block of the stream.
// `out` is used as a buffer: its size is its capacity, `n_out` the number of symbols in it.
inline bool decode_block(Bits& in, SymBuf& out, size_t& n_out, bool known_start, bool text_only, size_t max_out, bool* final, Tables& T)
{
	static const uint16_t lbase[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
	static const uint8_t lext[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
	static const uint16_t dbase[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
	static const uint8_t dext[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
	*final = in.get(1) != 0;
	const uint32_t type = in.get(2);
	if (type == 3)
		return false;
	if (type == 0) {
		in.pos = (in.pos + 7) & ~7ull;
		const size_t b = (size_t)(in.pos >> 3);
		if (b + 4 > in.nbytes)
			return false;
		const uint32_t len = in.base[b] | (in.base[b + 1] << 8), nlen = in.base[b + 2] | (in.base[b + 3] << 8);
		if ((len ^ 0xffffu) != nlen || b + 4 + len > in.nbytes)
			return false;
		out.ensure(n_out + len + (out.cap < n_out + len ? out.cap / 2 : 0));
		for (uint32_t i = 0; i < len; ++i) {
			if (text_only && !plausible_text(in.base[b + 4 + i]))
				return false;
			out.p[n_out++] = in.base[b + 4 + i];
		}
		in.pos += (4ull + len) * 8;
		return true;
	}
	if (type == 1 ? !T.set_fixed() : !T.read_dynamic(in, text_only))
		return false;
	const uint32_t* const lit = T.lit.data();
	const uint32_t* const dist = T.dist.data();
	const size_t start = n_out;
	// raw output pointer with room for one more match; the buffer grows geometrically
	size_t n = start, cap = 0;
	uint16_t* o = nullptr;
	auto grow = [&] {
		if (out.cap < n + (1u << 16) + 300)
			out.ensure(n + (1u << 16) + 300 + out.cap / 2);
		o = out.p;
		cap = out.cap - 300;
	};
	grow();
	const uint8_t* const base = in.base;
	const size_t safe_bytes = in.nbytes >= 8 ? in.nbytes - 8 : 0;
	uint64_t pos = in.pos;
	bool good = false;
	for (;;) {
		if (n >= cap) {
			if (n - start > max_out)
				break;
			grow();
		}
		uint64_t w;
		const size_t byte = (size_t)(pos >> 3);
		if (byte <= safe_bytes && in.nbytes >= 8) {
			memcpy(&w, base + byte, 8);
			w >>= (pos & 7);
		} else {
			in.pos = pos;
			if (in.past_end())
				break;
			w = in.peek();
		}
		uint32_t e = lit[w & ((1u << Tables::kLit) - 1)];
		int used = 0;
		if (e & 0x80) {
			const uint32_t sb = e & 0x7f;
			e = lit[(e >> 8) + ((w >> Tables::kLit) & ((1u << sb) - 1))];
			used = Tables::kLit;
		}
		if (!e)
			break;
		used += (int)(e & 0x7f);
		const uint32_t sym = e >> 8;
		if (sym < 256) {
			if (text_only && !plausible_text(sym))
				break;
			o[n++] = (uint16_t)sym;
			pos += (uint64_t)used;
			continue;
		}
		if (sym == 256) {
			pos += (uint64_t)used;
			good = true;
			break;
		}
		if (sym > 285)
			break;
		w >>= used;
		const uint32_t li = sym - 257;
		const uint32_t len = lbase[li] + (uint32_t)(w & ((1u << lext[li]) - 1));
		w >>= lext[li];
		used += lext[li];
		uint32_t d = dist[w & ((1u << Tables::kDist) - 1)];
		if (d & 0x80) {
			const uint32_t sb = d & 0x7f;
			d = dist[(d >> 8) + ((w >> Tables::kDist) & ((1u << sb) - 1))];
			w >>= Tables::kDist;
			used += Tables::kDist;
		}
		if (!d)
			break;
		w >>= (d & 0x7f);
		used += (int)(d & 0x7f);
		const uint32_t ds = d >> 8;
		if (ds > 29)
			break;
		const size_t distance = dbase[ds] + (size_t)(w & ((1u << dext[ds]) - 1));
		used += dext[ds];
		pos += (uint64_t)used;
		if (distance > n) {
			// reaches into the window in front of this chunk
			if (known_start || distance - n > 32768)
				break;
			for (uint32_t i = 0; i < len; ++i) {
				const ptrdiff_t j = (ptrdiff_t)(n + i) - (ptrdiff_t)distance;
				o[n + i] = j >= 0 ? o[j] : (uint16_t)(256 + 32768 + j);
			}
		} else if (distance >= 4) {
			// four symbols at a time (may write up to 3 past the match: room is reserved)
			const uint16_t* s = o + n - distance;
			uint16_t* t = o + n;
			for (uint32_t i = 0; i < len; i += 4) {
				uint64_t v;
				memcpy(&v, s + i, 8);
				memcpy(t + i, &v, 8);
			}
		} else {
			const uint16_t* s = o + n - distance;
			for (uint32_t i = 0; i < len; ++i)
				o[n + i] = s[i];
		}
		n += len;
	}
	n_out = n;
	in.pos = pos;
	return good;
}

// A block boundary at or after bit `from` (before `to`): see the header.  0 if none was found.
inline uint64_t find_block_start(const uint8_t* base, size_t nbytes, uint64_t from, uint64_t to)
{
	Tables T;
	SymBuf scratch;
	size_t ns = 0;
	for (uint64_t p = from; p < to; ++p) {
		Bits in{ base, nbytes, p };
		const uint64_t w = in.peek();
		if ((w & 7) != 4) // BFINAL = 0, BTYPE = 2 (dynamic)
			continue;
		// cheap rejects before the tables are built: HLIT <= 29, HDIST <= 29
		if (((w >> 3) & 31) > 29 || ((w >> 8) & 31) > 29)
			continue;
		ns = 0;
		bool final = false;
		if (!decode_block(in, scratch, ns, false, true, 1u << 22, &final, T) || final || ns < 64)
			continue;
		// and the block behind it must hold up as well
		const size_t n1 = ns;
		if (!decode_block(in, scratch, ns, false, true, 1u << 22, &final, T) || ns == n1)
			continue;
		return p;
	}
	return 0;
}

} // namespace pinf

// The gzip member [in, in + n) decompressed by `threads` threads; pull-driven like FastInflate.
// symbols -> bytes: a symbol below 256 is the byte itself, anything else names a byte of the 32 KB window in front of
// the chunk.  Past the first stretch of a chunk nearly every symbol is a plain byte: 32 at a time when they all are.
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2"))) inline size_t translate_symbols_avx2(const uint16_t* s, size_t n, const uint8_t* w, uint8_t* o)
{
	size_t k = 0;
	const __m256i hi = _mm256_set1_epi16((short)0xFF00);
	for (; k + 32 <= n; k += 32) {
		const __m256i a = _mm256_loadu_si256((const __m256i*)(s + k));
		const __m256i b = _mm256_loadu_si256((const __m256i*)(s + k + 16));
		if (_mm256_testz_si256(_mm256_or_si256(a, b), hi)) {
			// packus works per 128-bit lane: put the four quarters back in order
			const __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi16(a, b), 0xD8);
			_mm256_storeu_si256((__m256i*)(o + k), p);
		} else {
			for (size_t i = k; i < k + 32; ++i)
				o[i] = s[i] < 256 ? (uint8_t)s[i] : w[s[i] - 256];
		}
	}
	return k;
}
#endif
inline void translate_symbols(const uint16_t* s, size_t n, const uint8_t* w, uint8_t* o)
{
	size_t k = 0;
#if defined(__x86_64__) && defined(__GNUC__)
	static const bool have_avx2 = __builtin_cpu_supports("avx2");
	if (have_avx2)
		k = translate_symbols_avx2(s, n, w, o);
#endif
	for (; k < n; ++k)
		o[k] = s[k] < 256 ? (uint8_t)s[k] : w[s[k] - 256];
}

class ParInflate
{
  public:
	ParInflate(const uint8_t* in, size_t n, int threads, size_t chunk_bytes = 1u << 19)
	  : m_in(in)
	  , m_n(n)
	  , m_threads(std::max(1, threads))
	  , m_chunk(std::max<size_t>(chunk_bytes, 1u << 16))
	{
		if (!parse_header(0))
			start_sequential_from_scratch();
		else
			m_producer = std::thread([this] { producer_loop(); });
	}
	~ParInflate()
	{
		{
			std::lock_guard<std::mutex> lk(m_mu);
			m_quit = true;
		}
		m_cv.notify_all();
		if (m_producer.joinable())
			m_producer.join();
	}
	ParInflate(const ParInflate&) = delete;
	ParInflate& operator=(const ParInflate&) = delete;
	bool ok() const { return m_read_err.empty() && (!m_read_seq || m_read_seq->ok()); }
	std::string error() const { return !m_read_err.empty() ? m_read_err : (m_read_seq ? m_read_seq->error() : std::string()); }
	size_t parallel_chunks() const { return m_par_chunks; }

	long read(char* dst, size_t n)
	{
		size_t got = 0;
		while (got < n) {
			if (m_cur) { // a finished wave: its chunks go out one after the other
				if (m_serve < m_cur->good) {
					const std::vector<uint8_t>& b = m_cur->chunks[m_serve].bytes;
					const size_t c = std::min(n - got, b.size() - m_rpos);
					memcpy(dst + got, b.data() + m_rpos, c);
					m_rpos += c;
					got += c;
					if (m_rpos == b.size()) {
						m_serve++;
						m_rpos = 0;
					}
					continue;
				}
				{ // hand the buffers back to the producer
					std::lock_guard<std::mutex> lk(m_mu);
					m_cur->ready = false;
				}
				m_cv.notify_all();
				m_cur = nullptr;
				m_cons ^= 1;
			}
			if (m_producer.joinable()) {
				std::unique_lock<std::mutex> lk(m_mu);
				m_cv.wait(lk, [&] { return m_out[m_cons].ready || m_producer_done; });
				if (m_out[m_cons].ready) {
					m_cur = &m_out[m_cons];
					m_serve = 0;
					m_rpos = 0;
					continue;
				}
			}
			// the producer has stopped: error, end of the member, or the sequential decoder takes over
			if (!m_taken_over) { // (its verdict crosses to the reading thread here, after m_producer_done was seen under the lock)
				m_read_err = std::move(m_err);
				m_read_seq = std::move(m_seq);
				m_taken_over = true;
			}
			if (!m_read_err.empty())
				break;
			if (m_read_seq) {
				const long r = m_read_seq->read(dst + got, n - got);
				if (r <= 0)
					break;
				got += (size_t)r;
				continue;
			}
			break;
		}
		return (long)got;
	}

  private:
	// the member header at byte `off`: on success the member's first block starts at m_pos
	bool parse_header(size_t off)
	{
		// a member without surprises: magic, deflate, no reserved flags; optional fields skipped
		if (off + 18 + 8 > m_n || m_in[off] != 0x1f || m_in[off + 1] != 0x8b || m_in[off + 2] != 8 || (m_in[off + 3] & 0xe0))
			return false;
		const int flg = m_in[off + 3];
		size_t p = off + 10;
		if (flg & 4) {
			if (p + 2 > m_n)
				return false;
			p += 2 + (size_t)(m_in[p] | (m_in[p + 1] << 8));
		}
		for (int f : { 8, 16 })
			if (flg & f) {
				while (p < m_n && m_in[p])
					++p;
				++p;
			}
		if (flg & 2)
			p += 2;
		if (p + 8 >= m_n)
			return false;
		m_pos = (uint64_t)p * 8;
		m_known_start = true;
		m_member_off = off;
		return true;
	}

	// the sequential decoder takes the member that starts at m_member_off, nothing of which has been delivered
	void start_sequential_from_scratch()
	{
		if (m_member_off)
			m_seq.reset(new FastInflate(m_in + m_member_off, m_n - m_member_off, FastInflate::AfterMember()));
		else
			m_seq.reset(new FastInflate(m_in, m_n));
	}
	// hands the rest of the stream to the sequential decoder at the block boundary m_pos
	void start_sequential_here()
	{
		if (m_known_start) {
			start_sequential_from_scratch();
			return;
		}
		m_seq.reset(new FastInflate(m_in, m_n, m_pos, m_window.data(), m_crc, m_out_total));
	}

	struct Chunk
	{
		uint64_t start = 0; // bit position of its first block
		uint64_t end = 0;   // of the block behind its last one: must be hit exactly (stop_at_any: first boundary >= end)
		bool stop_at_any = false, known_start = false, ok = false, final = false;
		pinf::SymBuf sym;          // buffer ...
		size_t n_sym = 0;          // ... and the number of symbols in it
		std::vector<uint8_t> bytes;
		uint32_t crc = 0;
	};

	// one wave: up to m_threads chunks found, decoded, resolved and queued for read()
	// runs job(i) for i in [0, n) on the decoder threads, handing indices out one at a time
	template <class F>
	void run_jobs(size_t n, F job)
	{
		std::atomic<size_t> next{ 0 };
		std::vector<std::thread> th;
		const size_t nt = std::min<size_t>((size_t)m_threads, n);
		for (size_t t = 0; t < nt; ++t)
			th.emplace_back([&] {
				for (size_t i; (i = next.fetch_add(1)) < n;)
					job(i);
			});
		for (auto& t : th)
			t.join();
	}

	struct WaveOut
	{
		std::vector<Chunk> chunks; // (objects and buffers are reused from wave to wave)
		size_t good = 0;           // number of chunks that held up, in order
		bool ready = false;        // filled, waiting for read()
	};

	// waves are produced ahead of read() on a thread of their own, into two alternating sets of buffers
	void producer_loop()
	{
		int slot = 0;
		for (;;) {
			{
				std::unique_lock<std::mutex> lk(m_mu);
				m_cv.wait(lk, [&] { return m_quit || !m_out[slot].ready; });
				if (m_quit)
					break;
			}
			wave(m_out[slot]);
			const bool stop = m_done || m_seq || !m_err.empty();
			{
				std::lock_guard<std::mutex> lk(m_mu);
				m_out[slot].ready = m_out[slot].good > 0;
				if (stop)
					m_producer_done = true;
			}
			m_cv.notify_all();
			if (stop)
				break;
			slot ^= 1;
		}
		std::lock_guard<std::mutex> lk(m_mu);
		m_producer_done = true;
		m_cv.notify_all();
	}

	void wave(WaveOut& W)
	{
		W.good = 0;
		const uint64_t data_end = (uint64_t)(m_n - 8) * 8; // the trailer starts here at the latest
		// 1. chunk starts: the first is the validated position m_pos, the others are hunted for in parallel
		// (three chunks per thread and wave, dealt out dynamically, so that a slow chunk does not idle the rest)
		const int want = 3 * m_threads;
		std::vector<uint64_t> starts((size_t)want, 0);
		starts[0] = m_pos;
		run_jobs((size_t)want - 1, [&](size_t k) {
			const size_t i = k + 1;
			const uint64_t from = ((m_pos >> 3) + (uint64_t)i * m_chunk) * 8;
			if (from + 64 < data_end)
				starts[i] = pinf::find_block_start(m_in, m_n, from, std::min(data_end, from + (uint64_t)m_chunk * 8));
		});
		// chunks = stretches between consecutive starts that were found (a missing one merges two chunks); the last
		// chunk of the wave simply stops at the first block boundary behind its share of the input
		std::vector<Chunk>& chunks = W.chunks;
		if (chunks.size() < (size_t)want)
			chunks.resize((size_t)want);
		size_t n_chunks = 0;
		for (int i = 0; i < want; ++i) {
			if (i > 0 && !starts[(size_t)i])
				continue;
			if (n_chunks)
				chunks[n_chunks - 1].end = starts[(size_t)i];
			Chunk& c = chunks[n_chunks++];
			c.start = starts[(size_t)i];
			c.known_start = i == 0 && m_known_start;
			c.stop_at_any = c.ok = c.final = false;
			c.n_sym = 0;
		}
		chunks[n_chunks - 1].end = ((m_pos >> 3) + (uint64_t)want * m_chunk) * 8;
		chunks[n_chunks - 1].stop_at_any = true;
		// 2. decode every chunk symbolically until its end position (or the final block)
		{
			run_jobs(n_chunks, [&](size_t ci) {
					Chunk& c = chunks[ci];
					pinf::Tables T;
					pinf::Bits in{ m_in, m_n, c.start };
					c.sym.ensure(m_chunk * 6);
					for (;;) {
						bool final = false;
						if (!pinf::decode_block(in, c.sym, c.n_sym, c.known_start, false, 1u << 26, &final, T))
							return;
						if (final) {
							c.final = true;
							c.end = in.pos;
							c.ok = true;
							return;
						}
						if (in.pos >= c.end) {
							c.ok = c.stop_at_any || in.pos == c.end; // overshoot: the next start was not a real block boundary
							c.end = in.pos;
							return;
						}
						if (c.n_sym > ((size_t)1 << 30))
							return;
					}
				});
		}
		// 3. resolve in order as far as the chunks hold up; 4. translate + CRC in parallel
		size_t good = 0;
		while (good < n_chunks && chunks[good].ok) {
			++good;
			if (chunks[good - 1].final)
				break;
		}
		if (good == 0) {
			start_sequential_here();
			return;
		}
		std::vector<std::vector<uint8_t>> windows(good + 1);
		windows[0] = m_window;
		for (size_t i = 0; i < good; ++i) {
			// window behind chunk i = last 32 K of (window in front of it + its output)
			const uint16_t* const s = chunks[i].sym.p;
			const std::vector<uint8_t>& w = windows[i];
			std::vector<uint8_t>& nw = windows[i + 1];
			nw.resize(32768);
			const size_t n = chunks[i].n_sym;
			for (size_t k = 0; k < 32768; ++k) {
				// position counted from the end: the byte k places before the end of the stream so far
				const size_t back = 32768 - k; // 1 .. 32768
				if (back <= n) {
					const uint16_t v = s[n - back];
					nw[k] = v < 256 ? (uint8_t)v : w[v - 256];
				} else {
					nw[k] = w[32768 - (back - n)];
				}
			}
		}
		{
			run_jobs(good, [&](size_t i) {
					Chunk& c = chunks[i];
					const uint8_t* w = windows[i].data();
					c.bytes.resize(c.n_sym);
					const uint16_t* s = c.sym.p;
					uint8_t* o = c.bytes.data();
					translate_symbols(s, c.n_sym, w, o);
					c.crc = (uint32_t)crc32_z(0, o, c.bytes.size());
				});
		}
		W.good = good;
		for (size_t i = 0; i < good; ++i) {
			Chunk& c = chunks[i];
			m_crc = (uint32_t)crc32_combine(m_crc, c.crc, (z_off_t)c.bytes.size());
			m_out_total += c.bytes.size();
			m_pos = c.end;
			m_par_chunks++;
			if (c.final) {
				W.good = i + 1;
				finish_member();
				return;
			}
		}
		m_window = windows[good];
		m_known_start = false;
		if (good < n_chunks)
			start_sequential_here(); // a chunk did not end where the next one was thought to start
	}

	void finish_member()
	{
		m_done = true;
		const size_t t = (size_t)((m_pos + 7) >> 3);
		if (t + 8 > m_n) {
			m_err = "truncated gzip trailer";
			return;
		}
		const uint32_t crc = m_in[t] | (m_in[t + 1] << 8) | (m_in[t + 2] << 16) | ((uint32_t)m_in[t + 3] << 24);
		const uint32_t isize = m_in[t + 4] | (m_in[t + 5] << 8) | (m_in[t + 6] << 16) | ((uint32_t)m_in[t + 7] << 24);
		if (crc != m_crc)
			m_err = "incorrect data check (CRC-32)";
		else if (isize != (uint32_t)m_out_total)
			m_err = "incorrect length check";
		else if (t + 8 < m_n) {
			// more follows.  Another member of some size (files made with `cat a.gz b.gz`) is decoded like the first
			// one; short members, padding and anything else are the sequential decoder's business
			m_done = false;
			if (m_n - (t + 8) >= (m_chunk << 2) && parse_header(t + 8)) {
				m_crc = 0;
				m_out_total = 0;
				std::fill(m_window.begin(), m_window.end(), 0);
				return;
			}
			m_seq.reset(new FastInflate(m_in + t + 8, m_n - t - 8, FastInflate::AfterMember()));
		}
	}

	const uint8_t* m_in;
	size_t m_n;
	int m_threads;
	size_t m_chunk;
	uint64_t m_pos = 0; // bit position of the next block (always a validated block boundary)
	size_t m_member_off = 0; // byte offset of the member being decoded
	bool m_known_start = false, m_done = false;
	std::vector<uint8_t> m_window = std::vector<uint8_t>(32768, 0);
	uint32_t m_crc = 0;
	uint64_t m_out_total = 0;
	WaveOut m_out[2];
	WaveOut* m_cur = nullptr; // the wave read() is handing out
	int m_cons = 0;           // the set read() takes next
	size_t m_serve = 0, m_rpos = 0, m_par_chunks = 0;
	std::thread m_producer;
	std::mutex m_mu;
	std::condition_variable m_cv;
	bool m_quit = false, m_producer_done = false;
	std::unique_ptr<FastInflate> m_seq; // (the producer's; read() and ok() see m_read_seq / m_read_err)
	std::string m_err;
	std::unique_ptr<FastInflate> m_read_seq;
	std::string m_read_err;
	bool m_taken_over = false;
};

} // namespace arks_host
