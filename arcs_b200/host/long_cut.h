// long_cut.h -- the cutting rule of the reference's `long-to-linked-pe` (src/long-to-linked-pe.cpp:
// 233-292): a long read of at least max(2 l, m) bases becomes one pseudo-barcode (BX:Z:<1-based record
// number>); its sequence is cut into consecutive steps of 2 l bases, each a pseudo read pair (first l
// bases forward, next l bases reverse-complemented), plus one shorter pair for the remainder.
//
// Used by two callers: bin/long-to-linked-pe (the drop-in tool, long_to_linked_pe.cpp) and `arcs --arks
// --cut L` (LongCutSource below), which cuts the long reads while it reads them so that the arks-long
// pipeline (bin/arcs-make:300-312: long-to-linked-pe | arcs /dev/stdin) needs neither the second process,
// nor the pipe, nor the separate --bx-only pass for the multiplicities.  LongCutSource delivers the very
// bytes the tool writes, so everything behind it (parser, barcode rules, counters) is the code that a pipe
// would feed.
#pragma once
#include "seq_reader.h"

#include <string>

namespace arks_host {

// complement table: nucleotides and IUPAC ambiguity codes, case preserved, everything else unchanged
// (btllib::reverse_complement's behaviour, pinned by the golden of Examples/arks-long_test-demo)
struct ComplementTable
{
	unsigned char t[256];
	ComplementTable()
	{
		for (int i = 0; i < 256; ++i)
			t[i] = (unsigned char)i;
		const char* from = "ACGTURYSWKMBDHVN";
		const char* to = "TGCAAYRSWMKVHDBN";
		for (int i = 0; from[i]; ++i) {
			t[(unsigned char)from[i]] = (unsigned char)to[i];
			t[(unsigned char)(from[i] | 0x20)] = (unsigned char)(to[i] | 0x20);
		}
	}
};

struct LongCutter
{
	size_t l = 0;       // pseudo read length (-l)
	size_t m = 2000;    // minimum long-read length (-m)
	bool fasta = false; // '>' records without qualities (--fasta)

	bool accepts(size_t seq_size) const { return !(2 * l > seq_size || m > seq_size); }

	// reads of the barcode in the output: what --bx / --bx-only write (:236-246)
	size_t multiplicity(size_t seq_size) const
	{
		const size_t step = 2 * l;
		return seq_size % step != 0 ? (seq_size / step + 1) * 2 : seq_size / l;
	}

	// appends the pseudo pairs of record number `num` (0-based) to `o`; the caller has checked accepts()
	void append_pairs(std::string& o, const SeqRecord& record, size_t num) const
	{
		static const ComplementTable comp;
		const std::string& seq = record.seq;
		const std::string& qual = record.qual;
		const size_t seq_size = seq.size(), qual_size = qual.size();
		const size_t step = 2 * l;
		const char header_symbol = fasta ? '>' : '@';
		const std::string bx = " BX:Z:" + std::to_string(num + 1) + "\n";
		auto emit_header = [&](int read_num) {
			o.push_back(header_symbol);
			o += record.name;
			o += "_f";
			o += std::to_string(read_num);
			o += bx;
		};
		// one pseudo pair: forward piece [fpos, fpos + n), reverse-complemented piece [rpos, rpos + n)
		auto emit_pair = [&](int read_num, size_t fpos, size_t rpos, size_t n) {
			emit_header(read_num);
			o.append(seq, fpos, n);
			o.push_back('\n');
			if (!fasta) {
				o += "+\n";
				if (qual_size == 0)
					o.append(n, '#');
				else
					o.append(qual, fpos, std::min(n, qual_size > fpos ? qual_size - fpos : 0));
				o.push_back('\n');
			}
			emit_header(read_num);
			for (size_t i = 0; i < n; ++i)
				o.push_back((char)comp.t[(unsigned char)seq[rpos + n - 1 - i]]);
			o.push_back('\n');
			if (!fasta) {
				o += "+\n";
				if (qual_size == 0)
					o.append(n, '#');
				else
					for (size_t i = 0; i < n; ++i)
						o.push_back(qual[rpos + n - 1 - i]);
				o.push_back('\n');
			}
		};
		int read_num = 1;
		for (size_t i = 0; i <= seq_size - step; i += step) {
			emit_pair(read_num, i, i + l, l);
			++read_num;
		}
		const size_t remainder = seq_size % step;
		if (remainder != 0) {
			const size_t curr_i = seq_size - remainder;
			const size_t n = std::min(l, remainder); // seq.substr(curr_i, l).size()
			emit_pair(read_num, curr_i, seq_size - n, n);
		}
	}
};

// The output of `long-to-linked-pe -l L -m M <path>` as a byte stream, produced on demand.
class LongCutSource : public ByteSource
{
  public:
	LongCutSource(const std::string& path, size_t l, size_t m)
	  : m_reader(path, 1u << 22)
	{
		m_cut.l = l;
		m_cut.m = m;
	}
	bool ok() const { return m_reader.ok(); }
	long read(char* dst, size_t n) override
	{
		size_t got = 0;
		while (got < n) {
			if (m_pos == m_buf.size()) {
				m_buf.clear();
				m_pos = 0;
				// about a megabyte per refill; a record that is too short yields nothing
				while (!m_eof && m_buf.size() < (1u << 20)) {
					if (m_reader.read(m_record) < 0) {
						m_eof = true;
						break;
					}
					if (m_cut.accepts(m_record.seq.size()))
						m_cut.append_pairs(m_buf, m_record, m_num);
					++m_num;
				}
				if (m_buf.empty())
					break;
			}
			const size_t c = std::min(n - got, m_buf.size() - m_pos);
			memcpy(dst + got, m_buf.data() + m_pos, c);
			m_pos += c;
			got += c;
		}
		return (long)got;
	}

  private:
	SeqReader m_reader;
	LongCutter m_cut;
	SeqRecord m_record;
	std::string m_buf;
	size_t m_pos = 0;
	size_t m_num = 0; // 0-based record index within the file (btllib's record.num)
	bool m_eof = false;
};

} // namespace arks_host
