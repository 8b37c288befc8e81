// fasta_fast.h -- the draft (-f) read with every core.
//
// getContigKmers reads the draft through kseq one record at a time (Arcs/Arcs.cpp:1043-1094); a human-scale draft is
// 3 GB of text, and one thread walks it at well under 1 GB/s.  A plain (uncompressed) FASTA file whose shape is
// strict -- first byte '>', no CR, no NUL, no sequence line that starts with '>', '@' or '+' -- means the same to a
// line-based parser as to the reference's reader grammar (Arcs/kseq.h:175-215: name up to the first white space,
// sequence = the following lines up to the next line that starts with '>', '@' or '+', empty lines skipped).  Such a
// file is mapped, cut into byte ranges, and each thread lists the records that start in its range (header, body
// extent, sequence length); the caller decides what to keep and the threads then copy the bases it asks for.
// Anything else -- gzip, pipes, FASTQ drafts, odd characters -- returns false and the caller falls back to SeqReader.
#pragma once
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

namespace arks_host {

struct FastaRecord
{
	const char* name = nullptr;
	size_t name_n = 0;
	const char* body = nullptr; // the sequence lines (with their newlines)
	size_t body_n = 0;
	size_t seq_len = 0;         // body_n minus newlines
};

class MappedFasta
{
  public:
	MappedFasta() = default;
	MappedFasta(const MappedFasta&) = delete;
	MappedFasta& operator=(const MappedFasta&) = delete;
	~MappedFasta()
	{
		if (m_map)
			munmap((void*)m_map, m_size);
		if (m_fd >= 0)
			::close(m_fd);
	}

	// false: not a plain regular file in the strict shape (nothing is kept)
	bool open(const std::string& path, int threads)
	{
		m_fd = ::open(path.c_str(), O_RDONLY);
		if (m_fd < 0)
			return false;
		struct stat st;
		if (fstat(m_fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 2)
			return false;
		m_size = (size_t)st.st_size;
		void* m = mmap(nullptr, m_size, PROT_READ, MAP_PRIVATE, m_fd, 0);
		if (m == MAP_FAILED)
			return false;
		m_map = (const char*)m;
		if (m_map[0] != '>')
			return false;
		const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, threads), m_size / (4u << 20) + 1));
		std::vector<std::vector<FastaRecord>> parts((size_t)nt);
		std::vector<char> ok((size_t)nt, 1);
		auto work = [&](int t) {
			const size_t a = m_size * (size_t)t / (size_t)nt, b = m_size * (size_t)(t + 1) / (size_t)nt;
			ok[(size_t)t] = scan_range(a, b, parts[(size_t)t]) ? 1 : 0;
		};
		std::vector<std::thread> th;
		for (int t = 1; t < nt; ++t)
			th.emplace_back(work, t);
		work(0);
		for (auto& x : th)
			x.join();
		for (char o : ok)
			if (!o)
				return false;
		size_t n = 0;
		for (auto& p : parts)
			n += p.size();
		m_records.reserve(n);
		for (auto& p : parts)
			m_records.insert(m_records.end(), p.begin(), p.end());
		return true;
	}

	const std::vector<FastaRecord>& records() const { return m_records; }

	// bases [from, from + n) of a record's sequence (newlines skipped) -> dst
	static void copy_bases(const FastaRecord& r, size_t from, size_t n, char* dst)
	{
		if (r.seq_len + 1 >= r.body_n) { // a single line (with or without its newline)
			memcpy(dst, r.body + from, n);
			return;
		}
		const char* p = r.body;
		const char* const end = r.body + r.body_n;
		size_t skip = from;
		while (p < end && n) {
			const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
			const size_t len = (size_t)((nl ? nl : end) - p);
			if (skip >= len) {
				skip -= len;
			} else {
				const size_t take = std::min(n, len - skip);
				memcpy(dst, p + skip, take);
				dst += take;
				n -= take;
				skip = 0;
			}
			p = nl ? nl + 1 : end;
		}
	}

  private:
	// the records whose '>' lies in [a, b)
	bool scan_range(size_t a, size_t b, std::vector<FastaRecord>& out) const
	{
		const char* const base = m_map;
		const char* const fend = m_map + m_size;
		// first header at or after a
		const char* p = base + a;
		while (p < base + b) {
			if (*p == '>' && (p == base || p[-1] == '\n'))
				break;
			const char* nl = (const char*)memchr(p, '\n', (size_t)(fend - p));
			if (!nl)
				return true; // no line starts in the rest of the file
			p = nl + 1;
		}
		while (p < base + b && p < fend) {
			// header line
			const char* nl = (const char*)memchr(p, '\n', (size_t)(fend - p));
			const char* hend = nl ? nl : fend;
			if (memchr(p, '\r', (size_t)(hend - p)) || memchr(p, 0, (size_t)(hend - p)))
				return false;
			FastaRecord r;
			r.name = p + 1;
			const char* q = r.name;
			while (q < hend && !isspace((unsigned char)*q))
				++q;
			r.name_n = (size_t)(q - r.name);
			r.body = nl ? nl + 1 : fend;
			// body: lines up to the next one that starts with '>'
			const char* s = r.body;
			size_t newlines = 0;
			while (s < fend) {
				const char c = *s;
				if (c == '>')
					break;
				if (c == '@' || c == '+')
					return false; // the reference's reader would switch record type here
				const char* e = (const char*)memchr(s, '\n', (size_t)(fend - s));
				if (!e) {
					s = fend;
					break;
				}
				newlines++;
				s = e + 1;
			}
			r.body_n = (size_t)(s - r.body);
			if (memchr(r.body, '\r', r.body_n) || memchr(r.body, 0, r.body_n))
				return false;
			r.seq_len = r.body_n - newlines;
			out.push_back(r);
			p = s;
		}
		return true;
	}

	int m_fd = -1;
	const char* m_map = nullptr;
	size_t m_size = 0;
	std::vector<FastaRecord> m_records;
};

} // namespace arks_host
