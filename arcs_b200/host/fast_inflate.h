// fast_inflate.h -- gzip (RFC 1952) / DEFLATE (RFC 1951) decoder for the read ingest.
//
// The `arcs --arks` drop-in parses plain FASTQ at several GB/s (ingest.h), so for the usual .fq.gz
// input the single zlib inflate stream (about 0.3 GB/s of output) is what the whole run waits for.
// This decoder does the same job two to three times faster on FASTQ: 64-bit bit buffer refilled
// with one unaligned load, 11-bit primary Huffman tables with sub-tables, word-wise match copies,
// multi-member files, CRC-32 and ISIZE of every member verified (zlib's crc32, on a helper thread that works on
// one output window while the decoder fills the other).  It reads from memory (the memory-mapped compressed
// file) and is pull-driven: read(dst, n).
//
// Written from the two RFCs.  tests/test_host_cpu.py compares it byte for byte with zlib on streams
// of every block type, compression level and strategy, on multi-member files, header flags,
// truncation and corruption.  On malformed input it stops with an error message (ok() turns false)
// after handing out everything that was decoded before the defect, as gzread does.
#pragma once
#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h> // crc32() only

namespace arks_host {

class FastInflate
{
  public:
	FastInflate(const uint8_t* in, size_t n)
	  : m_ip(in)
	  , m_in_end(in + n)
	{
		for (auto& w : m_wins)
			w.resize(kHist + kOutCap + kSlack);
		m_crc_thread = std::thread([this] { crc_loop(); });
	}
	// continues a member in the middle: `bitpos` is the position of a block header, `window` the 32 KB of
	// output in front of it, `crc` / `member_out` the CRC-32 and length of the member's output so far
	FastInflate(const uint8_t* in, size_t n, uint64_t bitpos, const uint8_t* window, uint32_t crc, uint64_t member_out)
	  : m_ip(in + (size_t)(bitpos >> 3))
	  , m_in_end(in + n)
	{
		for (auto& w : m_wins)
			w.resize(kHist + kOutCap + kSlack);
		memcpy(m_wins[0].data(), window, kHist);
		refill();
		bits((int)(bitpos & 7));
		m_state = S_BLOCK_HEADER;
		m_total_out = kHist; // the whole window may be referenced
		m_crc = crc;
		m_member_out = member_out;
		m_members = 1;
		m_crc_thread = std::thread([this] { crc_loop(); });
	}
	// what follows a complete member: further members, or trailing bytes that are ignored
	struct AfterMember
	{
	};
	FastInflate(const uint8_t* in, size_t n, AfterMember)
	  : FastInflate(in, n)
	{
		m_members = 1;
	}
	~FastInflate()
	{
		{
			std::lock_guard<std::mutex> lk(m_mu);
			m_quit = true;
		}
		m_cv.notify_all();
		m_crc_thread.join();
	}
	FastInflate(const FastInflate&) = delete;
	FastInflate& operator=(const FastInflate&) = delete;

	bool ok() const { return m_err.empty(); }
	const std::string& error() const { return m_err; }

	// up to n decompressed bytes into dst; 0 = end of the stream (or error: see ok())
	long read(char* dst, size_t n)
	{
		size_t got = 0;
		while (got < n) {
			if (m_rpos == m_wpos) {
				if (m_state == S_DONE || !m_err.empty())
					break;
				produce();
				if (m_rpos == m_wpos && (m_state == S_DONE || !m_err.empty()))
					break;
				continue;
			}
			const size_t c = std::min(n - got, m_wpos - m_rpos);
			memcpy(dst + got, m_wins[m_cur].data() + m_rpos, c);
			m_rpos += c;
			got += c;
		}
		return (long)got;
	}

  private:
	static constexpr size_t kHist = 32768, kOutCap = 1u << 20, kSlack = 320;
	static constexpr int kLitBits = 11, kDistBits = 8;
	enum State { S_HEADER, S_BLOCK_HEADER, S_STORED, S_HUFFMAN, S_TRAILER, S_DONE };

	// ---- bit reader (LSB first)
	void refill()
	{
		if (m_in_end - m_ip >= 8) {
			uint64_t w;
			memcpy(&w, m_ip, 8);
			m_bitbuf |= w << m_bitcnt;
			m_ip += (63 - m_bitcnt) >> 3;
			m_bitcnt |= 56;
		} else {
			while (m_bitcnt <= 56) {
				if (m_ip < m_in_end) {
					m_bitbuf |= (uint64_t)*m_ip++ << m_bitcnt;
				} else {
					m_pad_bits += 8; // zeros past the end of the input; using them is an error
				}
				m_bitcnt += 8;
			}
		}
	}
	uint32_t bits(int n)
	{
		const uint32_t v = (uint32_t)(m_bitbuf & ((1ull << n) - 1));
		m_bitbuf >>= n;
		m_bitcnt -= n;
		return v;
	}
	bool overran() const { return m_pad_bits > 0 && m_bitcnt < (int)m_pad_bits; }
	// give the whole bytes still in the bit buffer back to the input (needs byte alignment)
	void unread_bytes()
	{
		const int real = m_bitcnt - (int)m_pad_bits;
		if (real > 0)
			m_ip -= real >> 3;
		m_bitbuf = 0;
		m_bitcnt = 0;
		m_pad_bits = 0;
	}

	bool fail(const char* what)
	{
		if (m_err.empty())
			m_err = what;
		return false;
	}

	// ---- canonical Huffman decode tables.  Entry: symbol << 8 | code length, or for a sub-table pointer
	// (offset << 8) | 0x80 | sub-table bits; 0 = no code.
	static uint32_t rev_bits(uint32_t c, int len)
	{
		uint32_t r = 0;
		for (int i = 0; i < len; ++i) {
			r = (r << 1) | (c & 1);
			c >>= 1;
		}
		return r;
	}
	bool build(const uint8_t* lens, int n, int primary, std::vector<uint32_t>& tab)
	{
		int count[16] = { 0 };
		for (int i = 0; i < n; ++i)
			count[lens[i]]++;
		count[0] = 0;
		int left = 1; // Kraft: codes still available
		for (int l = 1; l <= 15; ++l) {
			left = (left << 1) - count[l];
			if (left < 0)
				return fail("invalid Huffman code (over-subscribed)");
		}
		uint32_t next[16];
		uint32_t code = 0;
		for (int l = 1; l <= 15; ++l) {
			code = (code + count[l - 1]) << 1;
			next[l] = code;
		}
		tab.assign((size_t)1 << primary, 0);
		// longest code under every primary prefix that has long codes
		std::vector<uint8_t> sub_max((size_t)1 << primary, 0);
		std::vector<uint32_t> codes(n);
		for (int i = 0; i < n; ++i) {
			const int l = lens[i];
			if (!l)
				continue;
			codes[i] = rev_bits(next[l]++, l);
			if (l > primary) {
				uint8_t& m = sub_max[codes[i] & ((1u << primary) - 1)];
				if (l > m)
					m = (uint8_t)l;
			}
		}
		for (size_t p = 0; p < sub_max.size(); ++p)
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

	// Fast-loop table for literals/lengths, 2^kLitBits entries derived from m_lit: one lookup yields up to three
	// literals when their codes fit in the index (DNA lines code a base in ~2 bits).
	//   bits 0-3 total code bits, bits 4-5 number of literals - 1, bit 6 not a literal (symbol in bits 8-16: a
	//   length code, end of block), bit 7 go through m_lit (sub-table or invalid), bits 8-31 the literals
	void build_multi()
	{
		const uint32_t n = 1u << kLitBits;
		m_multi.assign(n, 0x80u);
		for (uint32_t i = 0; i < n; ++i) {
			const uint32_t e1 = m_lit[i];
			if (!e1 || (e1 & 0x80))
				continue;
			const uint32_t l1 = e1 & 0x7f, s1 = e1 >> 8;
			if (s1 >= 256) {
				m_multi[i] = l1 | 0x40u | (s1 << 8);
				continue;
			}
			uint32_t entry = l1 | (s1 << 8), used = l1, cnt = 1;
			while (cnt < 3) {
				const uint32_t e = m_lit[(i >> used) & (n - 1)]; // the bits above the index are unknown: only codes that fit count
				if (!e || (e & 0x80) || (e >> 8) >= 256 || used + (e & 0x7f) > (uint32_t)kLitBits)
					break;
				entry |= (e >> 8) << (8 + 8 * cnt);
				used += e & 0x7f;
				cnt++;
			}
			m_multi[i] = (entry & ~0xFu) | used | ((cnt - 1) << 4);
		}
	}

	// ---- CRC-32 of the output on a helper thread: one job per window
	struct CrcJob
	{
		int win = 0;
		size_t from = 0, to = 0;
		bool pending = false;
	};
	void crc_loop()
	{
		std::unique_lock<std::mutex> lk(m_mu);
		for (;;) {
			m_cv.wait(lk, [&] { return m_quit || m_jobs[m_job_tail].pending; });
			if (!m_jobs[m_job_tail].pending)
				return; // quit and nothing left
			const CrcJob j = m_jobs[m_job_tail];
			lk.unlock();
			const uint32_t c = (uint32_t)crc32(m_crc, m_wins[j.win].data() + j.from, (uInt)(j.to - j.from));
			lk.lock();
			m_crc = c;
			m_jobs[m_job_tail].pending = false;
			m_job_tail ^= 1;
			m_cv.notify_all();
		}
	}
	// hands the bytes [from, to) of window `win` to the helper (waits while that window's previous job runs)
	void crc_enqueue(int win, size_t from, size_t to)
	{
		if (to <= from)
			return;
		std::unique_lock<std::mutex> lk(m_mu);
		m_cv.wait(lk, [&] { return !m_jobs[m_job_head].pending; });
		m_jobs[m_job_head] = CrcJob{ win, from, to, true };
		m_job_head ^= 1;
		m_member_out += to - from;
		m_cv.notify_all();
	}
	void crc_wait_all()
	{
		std::unique_lock<std::mutex> lk(m_mu);
		m_cv.wait(lk, [&] { return !m_jobs[0].pending && !m_jobs[1].pending; });
	}

	// ---- the stream
	void produce()
	{
		// the decoder moves to the other window (the helper may still be reading this one), taking the last
		// kHist bytes along as history
		{
			const int next = m_cur ^ 1;
			{
				// a job on `next` is at most the one before the last: wait for it
				std::unique_lock<std::mutex> lk(m_mu);
				m_cv.wait(lk, [&] { return !(m_jobs[0].pending && m_jobs[0].win == next) && !(m_jobs[1].pending && m_jobs[1].win == next); });
			}
			const size_t keep = std::min(m_wpos, kHist);
			memcpy(m_wins[next].data() + kHist - keep, m_wins[m_cur].data() + m_wpos - keep, keep);
			m_cur = next;
			m_wpos = m_rpos = kHist;
		}
		const size_t round_start = m_wpos;
		size_t crc_from = m_wpos; // bytes from here on have not been handed to the helper yet
		auto fold_crc = [&] {
			crc_enqueue(m_cur, crc_from, m_wpos);
			crc_from = m_wpos;
		};
		while (m_err.empty() && m_state != S_DONE && m_wpos + 258 + 8 <= kHist + kOutCap) {
			switch (m_state) {
			case S_HEADER: header(); break;
			case S_BLOCK_HEADER: block_header(); break;
			case S_STORED: stored(); break;
			case S_HUFFMAN: huffman(); break;
			case S_TRAILER:
				fold_crc();
				crc_wait_all();
				trailer();
				break;
			default: break;
			}
			if (m_wpos - round_start >= kOutCap / 2)
				break;
		}
		fold_crc();
	}

	void header()
	{
		unread_bytes();
		if (m_ip == m_in_end) {
			m_state = S_DONE;
			return;
		}
		const uint8_t* p = m_ip;
		if (m_in_end - p < 18 || p[0] != 0x1f || p[1] != 0x8b) {
			if (m_members == 0)
				fail("not a gzip stream");
			m_state = S_DONE; // trailing garbage after a complete member is ignored, as zlib does
			return;
		}
		if (p[2] != 8 || (p[3] & 0xe0)) {
			fail("unsupported gzip header");
			return;
		}
		const int flg = p[3];
		p += 10;
		if (flg & 4) { // FEXTRA
			if (m_in_end - p < 2) {
				fail("truncated gzip header");
				return;
			}
			const size_t xlen = p[0] | (p[1] << 8);
			p += 2;
			if ((size_t)(m_in_end - p) < xlen) {
				fail("truncated gzip header");
				return;
			}
			p += xlen;
		}
		for (int f : { 8, 16 }) // FNAME, FCOMMENT
			if (flg & f) {
				while (p < m_in_end && *p)
					++p;
				if (p == m_in_end) {
					fail("truncated gzip header");
					return;
				}
				++p;
			}
		if (flg & 2) { // FHCRC
			if (m_in_end - p < 2) {
				fail("truncated gzip header");
				return;
			}
			p += 2;
		}
		m_ip = p;
		m_crc = 0;
		m_member_out = 0;
		m_members++;
		m_total_out = 0;
		m_state = S_BLOCK_HEADER;
	}

	void block_header()
	{
		refill();
		m_final = bits(1) != 0;
		const uint32_t type = bits(2);
		if (overran()) {
			fail("truncated deflate stream");
			return;
		}
		if (type == 0) {
			bits(m_bitcnt & 7); // to the byte boundary
			unread_bytes();
			if (m_in_end - m_ip < 4) {
				fail("truncated stored block");
				return;
			}
			const uint32_t len = m_ip[0] | (m_ip[1] << 8), nlen = m_ip[2] | (m_ip[3] << 8);
			if ((len ^ 0xffffu) != nlen) {
				fail("invalid stored block lengths");
				return;
			}
			m_ip += 4;
			m_stored_left = len;
			m_state = S_STORED;
		} else if (type == 1) {
			uint8_t l[288];
			for (int i = 0; i < 144; ++i)
				l[i] = 8;
			for (int i = 144; i < 256; ++i)
				l[i] = 9;
			for (int i = 256; i < 280; ++i)
				l[i] = 7;
			for (int i = 280; i < 288; ++i)
				l[i] = 8;
			uint8_t d[30];
			for (int i = 0; i < 30; ++i)
				d[i] = 5;
			if (build(l, 288, kLitBits, m_lit) && build(d, 30, kDistBits, m_dist)) {
				build_multi();
				m_state = S_HUFFMAN;
			}
		} else if (type == 2) {
			dynamic_tables();
		} else {
			fail("invalid block type");
		}
	}

	void dynamic_tables()
	{
		refill();
		const int hlit = (int)bits(5) + 257, hdist = (int)bits(5) + 1, hclen = (int)bits(4) + 4;
		if (hlit > 286 || hdist > 30) {
			fail("too many length or distance symbols");
			return;
		}
		static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
		uint8_t cl[19] = { 0 };
		for (int i = 0; i < hclen; ++i) {
			if (m_bitcnt < 3)
				refill();
			cl[order[i]] = (uint8_t)bits(3);
		}
		std::vector<uint32_t> cltab;
		if (!build(cl, 19, 7, cltab))
			return;
		uint8_t lens[286 + 30] = { 0 };
		int i = 0;
		while (i < hlit + hdist) {
			refill();
			const uint32_t e = cltab[m_bitbuf & 127];
			if (!e || (e & 0x80)) {
				fail("invalid code lengths set");
				return;
			}
			bits((int)(e & 0x7f));
			const int sym = (int)(e >> 8);
			if (sym < 16) {
				lens[i++] = (uint8_t)sym;
			} else {
				int rep, val = 0;
				if (sym == 16) {
					if (i == 0) {
						fail("invalid bit length repeat");
						return;
					}
					val = lens[i - 1];
					rep = 3 + (int)bits(2);
				} else if (sym == 17) {
					rep = 3 + (int)bits(3);
				} else {
					rep = 11 + (int)bits(7);
				}
				if (i + rep > hlit + hdist) {
					fail("invalid bit length repeat");
					return;
				}
				while (rep--)
					lens[i++] = (uint8_t)val;
			}
			if (overran()) {
				fail("truncated deflate stream");
				return;
			}
		}
		if (lens[256] == 0) {
			fail("invalid code -- missing end-of-block");
			return;
		}
		if (build(lens, hlit, kLitBits, m_lit) && build(lens + hlit, hdist, kDistBits, m_dist)) {
			build_multi();
			m_state = S_HUFFMAN;
		}
	}

	void stored()
	{
		const size_t room = kHist + kOutCap - m_wpos;
		const size_t c = std::min({ (size_t)m_stored_left, room, (size_t)(m_in_end - m_ip) });
		memcpy(m_wins[m_cur].data() + m_wpos, m_ip, c);
		m_ip += c;
		m_wpos += c;
		m_total_out += c;
		m_stored_left -= (uint32_t)c;
		if (m_stored_left == 0)
			m_state = m_final ? S_TRAILER : S_BLOCK_HEADER;
		else if (m_ip == m_in_end)
			fail("truncated stored block");
	}

	void huffman()
	{
		static const uint16_t lbase[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
		static const uint8_t lext[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
		static const uint16_t dbase[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
		static const uint8_t dext[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
		uint8_t* const out = m_wins[m_cur].data();
		size_t wpos = m_wpos;
		const size_t limit = kHist + kOutCap - 258 - 8;
		const uint32_t* const lit = m_lit.data();
		const uint32_t* const dist = m_dist.data();
		const uint32_t* const multi = m_multi.data();
		const size_t wstart = wpos;
		// the hot loop keeps the bit reader in locals; near the end of the input (or for anything unusual)
		// it falls back to the member-based reader below
		uint64_t bb = m_bitbuf;
		int bc = m_bitcnt;
		const uint8_t* ip = m_ip;
		const uint8_t* const fast_end = m_in_end - 8;
		bool fast = m_pad_bits == 0 && m_in_end - m_ip >= 8;
		const uint64_t hist_avail = m_total_out;
#define ARKS_REFILL()                                   \
	do {                                                \
		uint64_t w_;                                    \
		memcpy(&w_, ip, 8);                             \
		bb |= w_ << bc;                                 \
		ip += (63 - bc) >> 3;                           \
		bc |= 56;                                       \
	} while (0)
		while (fast && wpos <= limit && ip <= fast_end) {
			if (bc < 48)
				ARKS_REFILL();
			const uint32_t me = multi[bb & ((1u << kLitBits) - 1)];
			uint32_t sym;
			if (__builtin_expect((me & 0xC0u) == 0, 1)) {
				// one to three literals
				const uint32_t nb = me & 0xF;
				bb >>= nb;
				bc -= (int)nb;
				const uint32_t lits = me >> 8;
				memcpy(out + wpos, &lits, 4); // little endian: literal 1 first; up to 3 bytes too many (slack)
				wpos += ((me >> 4) & 3u) + 1u;
				continue;
			}
			if ((me & 0x80u) == 0) {
				bb >>= (me & 0xF);
				bc -= (int)(me & 0xF);
				sym = me >> 8;
			} else {
				uint32_t e = lit[bb & ((1u << kLitBits) - 1)];
				if (e & 0x80) {
					const uint32_t sb = e & 0x7f;
					e = lit[(e >> 8) + ((bb >> kLitBits) & ((1u << sb) - 1))];
					bb >>= kLitBits;
					bc -= kLitBits;
				}
				if (__builtin_expect(e == 0, 0)) {
					fail("invalid literal/length code");
					break;
				}
				bb >>= (e & 0x7f);
				bc -= (int)(e & 0x7f);
				sym = e >> 8;
				if (sym < 256) {
					out[wpos++] = (uint8_t)sym;
					continue;
				}
			}
			if (sym == 256) {
				m_state = m_final ? S_TRAILER : S_BLOCK_HEADER;
				break;
			}
			if (sym > 285) {
				fail("invalid literal/length code");
				break;
			}
			if (bc < 48)
				ARKS_REFILL();
			const uint32_t li = sym - 257;
			const uint32_t len = lbase[li] + (uint32_t)(bb & ((1u << lext[li]) - 1));
			bb >>= lext[li];
			bc -= lext[li];
			uint32_t d = dist[bb & ((1u << kDistBits) - 1)];
			if (__builtin_expect((d & 0x80) != 0, 0)) {
				const uint32_t sb = d & 0x7f;
				d = dist[(d >> 8) + ((bb >> kDistBits) & ((1u << sb) - 1))];
				bb >>= kDistBits;
				bc -= kDistBits;
			}
			if (__builtin_expect(d == 0, 0)) {
				fail("invalid distance code");
				break;
			}
			bb >>= (d & 0x7f);
			bc -= (int)(d & 0x7f);
			const uint32_t ds = d >> 8;
			if (ds > 29) {
				fail("invalid distance code");
				break;
			}
			const size_t distance = dbase[ds] + (size_t)(bb & ((1u << dext[ds]) - 1));
			bb >>= dext[ds];
			bc -= dext[ds];
			if (distance > hist_avail + (wpos - wstart) || distance > kHist) {
				fail("invalid distance too far back");
				break;
			}
			uint8_t* dst = out + wpos;
			const uint8_t* src = dst - distance;
			wpos += len;
			if (distance >= 8) {
				uint8_t* const end = dst + len;
				do {
					uint64_t w;
					memcpy(&w, src, 8);
					memcpy(dst, &w, 8);
					src += 8;
					dst += 8;
				} while (dst < end);
			} else {
				// the match overlaps itself: repeat its first `distance` bytes, eight at a time
				uint8_t pat[8];
				for (uint32_t i = 0; i < 8; ++i)
					pat[i] = src[i % distance];
				uint64_t w;
				memcpy(&w, pat, 8);
				const uint32_t stride = (uint32_t)(8 / distance * distance);
				uint8_t* const end = dst + len;
				do {
					memcpy(dst, &w, 8);
					dst += stride;
				} while (dst < end);
			}
		}
#undef ARKS_REFILL
		m_bitbuf = bb;
		m_bitcnt = bc;
		m_ip = ip;
		// the careful loop: last bytes of the input, and the symbol after the fast loop stopped
		while (m_err.empty() && m_state == S_HUFFMAN && wpos <= limit) {
			if (m_bitcnt < 48)
				refill();
			uint32_t e = lit[m_bitbuf & ((1u << kLitBits) - 1)];
			if (e & 0x80) {
				const uint32_t sb = e & 0x7f;
				e = lit[(e >> 8) + ((m_bitbuf >> kLitBits) & ((1u << sb) - 1))];
				m_bitbuf >>= kLitBits;
				m_bitcnt -= kLitBits;
			}
			if (!e) {
				fail("invalid literal/length code");
				break;
			}
			m_bitbuf >>= (e & 0x7f);
			m_bitcnt -= (int)(e & 0x7f);
			const uint32_t sym = e >> 8;
			if (overran()) // the code came out of the zero padding behind a truncated input: nothing of it is real
				break;
			if (sym < 256) {
				out[wpos++] = (uint8_t)sym;
				if (m_in_end - m_ip >= 8 && m_pad_bits == 0)
					break; // back to the fast loop on the next call
				continue;
			}
			if (sym == 256) {
				m_state = m_final ? S_TRAILER : S_BLOCK_HEADER;
				break;
			}
			if (sym > 285) {
				fail("invalid literal/length code");
				break;
			}
			const uint32_t li = sym - 257;
			const uint32_t len = lbase[li] + bits(lext[li]);
			uint32_t d = dist[m_bitbuf & ((1u << kDistBits) - 1)];
			if (d & 0x80) {
				const uint32_t sb = d & 0x7f;
				d = dist[(d >> 8) + ((m_bitbuf >> kDistBits) & ((1u << sb) - 1))];
				m_bitbuf >>= kDistBits;
				m_bitcnt -= kDistBits;
			}
			if (!d) {
				fail("invalid distance code");
				break;
			}
			m_bitbuf >>= (d & 0x7f);
			m_bitcnt -= (int)(d & 0x7f);
			const uint32_t ds = d >> 8;
			if (ds > 29) {
				fail("invalid distance code");
				break;
			}
			const size_t distance = dbase[ds] + bits(dext[ds]);
			if (overran())
				break;
			if (distance > hist_avail + (wpos - wstart) || distance > kHist) {
				fail("invalid distance too far back");
				break;
			}
			uint8_t* dst = out + wpos;
			const uint8_t* src = dst - distance;
			wpos += len;
			for (uint32_t i = 0; i < len; ++i)
				dst[i] = src[i];
			if (m_in_end - m_ip >= 8 && m_pad_bits == 0)
				break;
		}
		if (overran())
			fail("truncated deflate stream");
		m_total_out += wpos - wstart;
		m_wpos = wpos;
	}

	void trailer()
	{
		bits(m_bitcnt & 7);
		unread_bytes();
		if (m_in_end - m_ip < 8) {
			fail("truncated gzip trailer");
			return;
		}
		const uint32_t crc = m_ip[0] | (m_ip[1] << 8) | (m_ip[2] << 16) | ((uint32_t)m_ip[3] << 24);
		const uint32_t isize = m_ip[4] | (m_ip[5] << 8) | (m_ip[6] << 16) | ((uint32_t)m_ip[7] << 24);
		m_ip += 8;
		if (crc != m_crc) {
			fail("incorrect data check (CRC-32)");
			return;
		}
		if (isize != (uint32_t)m_member_out) {
			fail("incorrect length check");
			return;
		}
		m_state = S_HEADER;
	}

	const uint8_t* m_ip;
	const uint8_t* const m_in_end;
	uint64_t m_bitbuf = 0;
	int m_bitcnt = 0;
	unsigned m_pad_bits = 0;
	std::vector<uint8_t> m_wins[2];
	int m_cur = 0;
	size_t m_wpos = kHist, m_rpos = kHist;
	std::mutex m_mu;
	std::condition_variable m_cv;
	std::thread m_crc_thread;
	CrcJob m_jobs[2];
	int m_job_head = 0, m_job_tail = 0;
	bool m_quit = false;
	State m_state = S_HEADER;
	bool m_final = false;
	uint32_t m_stored_left = 0;
	std::vector<uint32_t> m_lit, m_dist, m_multi;
	uint32_t m_crc = 0;
	uint64_t m_member_out = 0, m_total_out = 0;
	size_t m_members = 0;
	std::string m_err;
};

} // namespace arks_host
