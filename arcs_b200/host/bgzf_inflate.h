// bgzf_inflate.h -- multi-threaded decoding of BGZF input (bgzip, htslib): a gzip file made of many small
// members (at most 64 KB of data each) whose headers carry the member's compressed size in a "BC" extra
// sub-field.  Members do not share a window, so they are simply dealt out to threads: the producer walks the
// headers of a wave of members, sums the ISIZE trailers into output offsets, and every member is inflated
// straight into its place and checked against its CRC-32 and ISIZE.
//
// Pull-driven like FastInflate / ParInflate (read / ok / error).  A member that is not BGZF (a plain gzip member
// appended to the file, trailing bytes) hands the rest of the file to FastInflate; a member that does not
// decode ends the stream there with an error, the members before it are delivered.
//
// Part of the read ingest of the ARKS path (reference: gzopen/gzread under kseq, Arcs.cpp:1162-1170, which
// read such files one member after the other).
#pragma once
#include "fast_inflate.h"

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>

namespace arks_host {

class BgzfInflate
{
  public:
	struct Member
	{
		size_t pos = 0;      // of the member's first byte
		size_t data = 0;     // of its deflate stream
		size_t size = 0;     // bytes of the whole member (BSIZE + 1)
		uint32_t isize = 0;  // bytes it inflates to
		uint32_t crc = 0;    // of those bytes
		size_t out = 0;      // where they go in the wave's buffer
	};

	// Parses the member header at in[pos..): true if it is a BGZF member that lies completely inside the file.
	static bool parse_member(const uint8_t* in, size_t n, size_t pos, Member& m)
	{
		if (pos + 18 + 8 > n || in[pos] != 0x1f || in[pos + 1] != 0x8b || in[pos + 2] != 8 || in[pos + 3] != 4)
			return false; // BGZF sets FEXTRA and nothing else
		const size_t xlen = (size_t)in[pos + 10] | ((size_t)in[pos + 11] << 8);
		const size_t xend = pos + 12 + xlen;
		if (xend + 8 > n)
			return false;
		size_t bsize = 0;
		for (size_t p = pos + 12; p + 4 <= xend;) {
			const size_t slen = (size_t)in[p + 2] | ((size_t)in[p + 3] << 8);
			if (in[p] == 'B' && in[p + 1] == 'C' && slen == 2 && p + 6 <= xend)
				bsize = ((size_t)in[p + 4] | ((size_t)in[p + 5] << 8)) + 1;
			p += 4 + slen;
		}
		if (bsize < 12 + xlen + 8 || pos + bsize > n)
			return false;
		m.pos = pos;
		m.data = xend;
		m.size = bsize;
		const uint8_t* t = in + pos + bsize - 8;
		m.crc = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
		m.isize = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
		return m.isize <= (1u << 16); // the format's limit; anything else is not BGZF
	}
	static bool is_bgzf(const uint8_t* in, size_t n)
	{
		Member m;
		return parse_member(in, n, 0, m);
	}

	BgzfInflate(const uint8_t* in, size_t n, int threads, size_t wave_bytes_per_thread = 4u << 20)
	  : m_in(in)
	  , m_n(n)
	  , m_threads(std::max(1, threads))
	  , m_wave_bytes(std::max<size_t>(wave_bytes_per_thread, 1u << 16) * (size_t)std::max(1, threads))
	{
		m_producer = std::thread([this] { producer_loop(); });
	}
	~BgzfInflate()
	{
		{
			std::lock_guard<std::mutex> lk(m_mu);
			m_quit = true;
		}
		m_cv.notify_all();
		if (m_producer.joinable())
			m_producer.join();
	}
	BgzfInflate(const BgzfInflate&) = delete;
	BgzfInflate& operator=(const BgzfInflate&) = delete;
	bool ok() const { return m_err.empty() && (!m_seq || m_seq->ok()); }
	std::string error() const { return !m_err.empty() ? m_err : (m_seq ? m_seq->error() : std::string()); }
	size_t parallel_members() const { return m_par_members; }

	long read(char* dst, size_t n)
	{
		size_t got = 0;
		while (got < n) {
			if (m_cur) {
				if (m_rpos < m_cur->good_bytes) {
					const size_t c = std::min(n - got, m_cur->good_bytes - m_rpos);
					memcpy(dst + got, m_cur->buf.get() + m_rpos, c);
					m_rpos += c;
					got += c;
					continue;
				}
				{ // the buffer goes back to the producer
					std::lock_guard<std::mutex> lk(m_mu);
					m_cur->ready = false;
				}
				m_cv.notify_all();
				m_cur = nullptr;
				m_cons ^= 1;
			}
			{
				std::unique_lock<std::mutex> lk(m_mu);
				m_cv.wait(lk, [&] { return m_out[m_cons].ready || m_producer_done; });
				if (m_out[m_cons].ready) {
					m_cur = &m_out[m_cons];
					m_rpos = 0;
					continue;
				}
			}
			// the producer has stopped: end of the file, an error, or the sequential decoder takes over
			if (!m_taken_over) { // (its verdict crosses to this thread here, after m_producer_done was seen under the lock)
				m_err = std::move(m_producer_err);
				m_seq = std::move(m_producer_seq);
				m_taken_over = true;
			}
			if (!m_err.empty() || !m_seq)
				break;
			const long r = m_seq->read(dst + got, n - got);
			if (r <= 0)
				break;
			got += (size_t)r;
		}
		return (long)got;
	}

  private:
	struct Wave
	{
		std::vector<Member> members;
		std::unique_ptr<uint8_t[]> buf; // grown, never zero-filled
		size_t cap = 0;
		size_t good_bytes = 0; // of the members that held up, in order
		bool ready = false;
	};

	void producer_loop()
	{
		int slot = 0;
		bool more = true;
		while (more) {
			Wave& W = m_out[slot];
			{
				std::unique_lock<std::mutex> lk(m_mu);
				m_cv.wait(lk, [&] { return !W.ready || m_quit; });
				if (m_quit)
					break;
			}
			more = wave(W);
			if (W.good_bytes || !more) {
				{
					std::lock_guard<std::mutex> lk(m_mu);
					if (W.good_bytes)
						W.ready = true;
				}
				m_cv.notify_all();
				if (W.good_bytes)
					slot ^= 1;
			}
		}
		{
			std::lock_guard<std::mutex> lk(m_mu);
			m_producer_done = true;
		}
		m_cv.notify_all();
	}

	// Decodes the next members; false when nothing follows from this thread (end, error, hand-over).
	bool wave(Wave& W)
	{
		W.members.clear();
		W.good_bytes = 0;
		size_t out = 0;
		bool bgzf = true;
		while (m_pos < m_n && out < m_wave_bytes && W.members.size() < (1u << 16)) {
			Member m;
			if (!parse_member(m_in, m_n, m_pos, m)) {
				bgzf = false;
				break;
			}
			m.out = out;
			out += m.isize;
			m_pos += m.size;
			W.members.push_back(m);
		}
		if (out > W.cap) {
			W.buf.reset(new uint8_t[out]);
			W.cap = out;
		}
		const size_t nm = W.members.size();
		std::vector<uint8_t> bad(nm, 0);
		std::atomic<size_t> next{ 0 };
		auto work = [&] {
			z_stream z;
			memset(&z, 0, sizeof z);
			if (inflateInit2(&z, -15) != Z_OK) {
				for (size_t i; (i = next.fetch_add(1)) < nm;)
					bad[i] = 1;
				return;
			}
			for (size_t i; (i = next.fetch_add(1)) < nm;) {
				const Member& m = W.members[i];
				uint8_t* dst = W.buf.get() + m.out;
				uint8_t dummy;
				inflateReset(&z);
				z.next_in = const_cast<Bytef*>(m_in + m.data);
				z.avail_in = (uInt)(m.pos + m.size - 8 - m.data);
				z.next_out = m.isize ? dst : &dummy;
				z.avail_out = m.isize; // (a stream that wants to write more than ISIZE stops with Z_OK / Z_BUF_ERROR)
				const int r = inflate(&z, Z_FINISH);
				if (r != Z_STREAM_END || z.total_out != m.isize || z.avail_in != 0 ||
				    (uint32_t)crc32_z(0, dst, m.isize) != m.crc)
					bad[i] = 1;
			}
			inflateEnd(&z);
		};
		{
			std::vector<std::thread> th;
			const size_t nt = std::min<size_t>((size_t)m_threads, (nm + 7) / 8);
			for (size_t t = 1; t < nt; ++t)
				th.emplace_back(work);
			if (nm)
				work();
			for (auto& t : th)
				t.join();
		}
		size_t good = 0;
		while (good < nm && !bad[good])
			++good;
		W.good_bytes = good == nm ? out : W.members[good].out;
		m_par_members += good;
		if (good < nm) {
			m_producer_err = "bgzf member at byte " + std::to_string(W.members[good].pos) + " does not decode (deflate data, CRC-32 or size)";
			return false;
		}
		if (!bgzf) {
			// something else follows (a plain gzip member, padding, a cut-off member): the sequential decoder's business
			m_producer_seq.reset(new FastInflate(m_in + m_pos, m_n - m_pos));
			return false;
		}
		return m_pos < m_n;
	}

	const uint8_t* m_in;
	size_t m_n;
	int m_threads;
	size_t m_wave_bytes;
	size_t m_pos = 0;
	size_t m_par_members = 0;
	std::string m_err, m_producer_err;
	std::unique_ptr<FastInflate> m_seq, m_producer_seq;
	bool m_taken_over = false;

	Wave m_out[2];
	Wave* m_cur = nullptr;
	int m_cons = 0;
	size_t m_rpos = 0;
	std::mutex m_mu;
	std::condition_variable m_cv;
	bool m_quit = false, m_producer_done = false;
	std::thread m_producer;
};

} // namespace arks_host
