// seq_reader.h -- streaming FASTA/FASTQ reader over zlib with the record grammar of the
// reader the reference uses (klib's kseq as vendored in Arcs/kseq.h:175-215), written from
// that grammar's description, not from its macros:
//
//  * between records (at start, or after a FASTQ record) skip bytes until the next '>' or '@';
//  * name = bytes up to the first isspace(); if that delimiter was not '\n' the rest of the
//    line is the comment; a trailing '\r' is dropped from a line that has more than one char;
//  * the sequence is every following line up to a line that STARTS with '>', '@' or '+';
//    empty lines are skipped;
//  * '+' starts the quality: skip that line, then append lines until quality is at least as
//    long as the sequence; a length mismatch is error -2, EOF is -1.
//
// Works for plain and gzip input, files and pipes (open_source: gzip files on disk through the fast decoder of
// fast_inflate.h, the rest through zlib, whose gzread passes uncompressed data through).
#pragma once
#include "bgzf_inflate.h"
#include "par_inflate.h"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <memory>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>
#include <zlib.h>

namespace arks_host {

// ---- where the bytes come from -------------------------------------------------------------------
struct ByteSource
{
	virtual ~ByteSource() {}
	virtual long read(char* dst, size_t n) = 0; // <= 0: end of the stream
};

// zlib: gzip or plain, files and pipes
struct ZlibSource : ByteSource
{
	gzFile fp;
	explicit ZlibSource(gzFile f)
	  : fp(f)
	{
	}
	~ZlibSource() override
	{
		if (fp)
			gzclose(fp);
	}
	long read(char* dst, size_t n) override { return fp ? gzread(fp, dst, (unsigned)std::min<size_t>(n, 1u << 30)) : 0; }
};

// a gzip file on disk (memory-mapped) through fast_inflate.h, or, when it is large and threads are to be had,
// through the multi-threaded decoders of par_inflate.h (one DEFLATE stream) or bgzf_inflate.h (bgzip's members)
struct FastGzSource : ByteSource
{
	std::string path;
	const uint8_t* map;
	size_t size;
	std::unique_ptr<FastInflate> inf;
	std::unique_ptr<ParInflate> par;
	std::unique_ptr<BgzfInflate> bgzf;
	bool warned = false;
	FastGzSource(const std::string& p, const uint8_t* m, size_t n, int threads)
	  : path(p)
	  , map(m)
	  , size(n)
	{
		if (threads > 1 && BgzfInflate::is_bgzf(m, n))
			bgzf.reset(new BgzfInflate(m, n, threads));
		else if (threads > 1)
			par.reset(new ParInflate(m, n, threads));
		else
			inf.reset(new FastInflate(m, n));
	}
	~FastGzSource() override
	{
		inf.reset();
		par.reset();
		bgzf.reset();
		munmap((void*)map, size);
	}
	long read(char* dst, size_t n) override
	{
		const long got = bgzf ? bgzf->read(dst, n) : par ? par->read(dst, n) : inf->read(dst, n);
		if (!(bgzf ? bgzf->ok() : par ? par->ok() : inf->ok()) && !warned) {
			// like a gzread error upstream, a damaged file ends the input where the damage is -- but not silently
			fprintf(stderr, "arcs: warning: %s: gzip stream ends early: %s\n", path.c_str(),
			    (bgzf ? bgzf->error() : par ? par->error() : inf->error()).c_str());
			warned = true;
		}
		return got;
	}
};

// decoder threads for a gzip file of `bytes` compressed bytes: ARKS_GZ_THREADS, else one per core (at most 32) for
// files of at least 8 MB (measured on 8 cores: 4 threads 0.51, 6 0.64, 8 0.91, 12 0.79 GB/s of text -- with gzip input
// the decoder is the bottleneck of the whole run and the parser threads behind it mostly sleep); small files are not
// worth the threads
inline int gz_threads_for(size_t bytes)
{
	if (const char* e = getenv("ARKS_GZ_THREADS"))
		return std::max(1, atoi(e));
	if (bytes < (8u << 20))
		return 1;
	const unsigned hw = std::thread::hardware_concurrency();
	return (int)std::min(32u, std::max(1u, hw));
}

// Opens `path` for reading: gzip files on disk get the fast decoder (ARKS_ZLIB=1: always zlib), everything
// else (plain files, pipes) goes through zlib, which passes plain data through.  Null if it cannot be opened.
inline std::unique_ptr<ByteSource> open_source(const std::string& path)
{
	const int fd = ::open(path.c_str(), O_RDONLY);
	if (fd < 0)
		return nullptr;
	struct stat st;
	unsigned char magic[2] = { 0, 0 };
	const bool regular = fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 18;
	if (regular && !getenv("ARKS_ZLIB") && pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
		void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
		if (m != MAP_FAILED) {
			madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
			::close(fd);
			return std::unique_ptr<ByteSource>(new FastGzSource(path, (const uint8_t*)m, (size_t)st.st_size, gz_threads_for((size_t)st.st_size)));
		}
	}
	gzFile fp = gzdopen(fd, "r");
	if (!fp) {
		::close(fd);
		return nullptr;
	}
	gzbuffer(fp, 1u << 18);
	return std::unique_ptr<ByteSource>(new ZlibSource(fp));
}

struct SeqRecord
{
	std::string name, comment, seq, qual;
	// The reference copies name / comment / sequence out of the reader as C strings
	// (`sequence = seq->seq.s`, Arcs.cpp:1053,1190-1200), so everything from the first NUL byte
	// on is dropped.  Call after read().
	void truncate_at_nul()
	{
		for (std::string* s : { &name, &comment, &seq }) {
			size_t z = s->find('\0');
			if (z != std::string::npos)
				s->resize(z);
		}
	}
};

class SeqReader
{
  public:
	explicit SeqReader(const std::string& path, size_t bufsize = 1u << 20)
	  : m_src(open_source(path))
	  , m_buf(bufsize)
	{
		m_opened = m_src != nullptr;
	}
	// continues an already open stream: `prefix` (bytes that were read ahead of it) comes first.
	// `src` may be null: then only the prefix is read.
	SeqReader(std::unique_ptr<ByteSource> src, std::string prefix, size_t bufsize = 1u << 20)
	  : m_src(std::move(src))
	  , m_buf(bufsize)
	  , m_prefix(std::move(prefix))
	  , m_opened(true)
	{
	}
	SeqReader(const SeqReader&) = delete;
	SeqReader& operator=(const SeqReader&) = delete;
	bool ok() const { return m_opened; }

	// >= 0: sequence length; -1: end of file; -2: truncated / mismatched quality
	int read(SeqRecord& r)
	{
		int c;
		if (m_last == 0) {
			while ((c = getc()) != -1 && c != '>' && c != '@') {
			}
			if (c == -1)
				return -1;
			m_last = c;
		}
		r.comment.clear();
		r.seq.clear();
		r.qual.clear();
		int delim = 0;
		if (!until_space(r.name, &delim))
			return -1;
		if (delim != '\n')
			line(r.comment, false);
		while ((c = getc()) != -1 && c != '>' && c != '+' && c != '@') {
			if (c == '\n')
				continue;
			r.seq.push_back((char)c);
			line(r.seq, true);
		}
		if (c == '>' || c == '@')
			m_last = c;
		if (c != '+')
			return (int)r.seq.size();
		while ((c = getc()) != -1 && c != '\n') {
		}
		if (c == -1)
			return -2;
		while (line(r.qual, true) && r.qual.size() < r.seq.size()) {
		}
		m_last = 0;
		if (r.seq.size() != r.qual.size())
			return -2;
		return (int)r.seq.size();
	}

  private:
	bool fill()
	{
		if (m_eof)
			return false;
		m_begin = 0;
		if (m_prefix_pos < m_prefix.size()) {
			const size_t n = std::min(m_buf.size(), m_prefix.size() - m_prefix_pos);
			memcpy(m_buf.data(), m_prefix.data() + m_prefix_pos, n);
			m_prefix_pos += n;
			m_end = n;
			if (m_prefix_pos == m_prefix.size())
				std::string().swap(m_prefix), m_prefix_pos = 0;
			return true;
		}
		const long n = m_src ? m_src->read(m_buf.data(), m_buf.size()) : 0;
		m_end = n > 0 ? (size_t)n : 0;
		if (m_end == 0) {
			m_eof = true;
			return false;
		}
		return true;
	}
	int getc()
	{
		if (m_begin >= m_end && !fill())
			return -1;
		return (unsigned char)m_buf[m_begin++];
	}
	// reads up to (not including) the first whitespace byte; false at EOF with nothing read
	bool until_space(std::string& s, int* delim)
	{
		s.clear();
		*delim = 0;
		bool got = false;
		for (;;) {
			if (m_begin >= m_end && !fill())
				break;
			got = true;
			size_t i = m_begin;
			while (i < m_end && !isspace((unsigned char)m_buf[i]))
				++i;
			s.append(m_buf.data() + m_begin, i - m_begin);
			m_begin = i + 1;
			if (i < m_end) {
				*delim = (unsigned char)m_buf[i];
				break;
			}
		}
		return got;
	}
	// appends the rest of the current line; false at EOF with nothing read.  As in the
	// reference's reader the '\r' test looks at the whole accumulated string.
	bool line(std::string& s, bool append)
	{
		if (!append)
			s.clear();
		bool got = false;
		for (;;) {
			if (m_begin >= m_end && !fill())
				break;
			got = true;
			const char* p = m_buf.data() + m_begin;
			const char* nl = (const char*)memchr(p, '\n', m_end - m_begin);
			size_t n = nl ? (size_t)(nl - p) : m_end - m_begin;
			s.append(p, n);
			m_begin += n + 1;
			if (nl)
				break;
		}
		if (!got)
			return false;
		if (s.size() > 1 && s.back() == '\r')
			s.pop_back();
		return true;
	}

	std::unique_ptr<ByteSource> m_src;
	std::vector<char> m_buf;
	std::string m_prefix;
	size_t m_prefix_pos = 0;
	bool m_opened = false;
	size_t m_begin = 0, m_end = 0;
	bool m_eof = false;
	int m_last = 0;
};

} // namespace arks_host
