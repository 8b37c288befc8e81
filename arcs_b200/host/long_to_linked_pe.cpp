// long_to_linked_pe.cpp -- drop-in for the reference's `long-to-linked-pe` (src/long-to-linked-pe.cpp:
// 66-325), the arks-long pre-processing step of bin/arcs-make:233,300-311: every long read of at
// least max(2 l, m) bases becomes one pseudo-barcode (BX:Z:<1-based record number>) whose sequence
// is cut into consecutive 2l-base steps; each step yields a pseudo read pair (first l bases forward,
// next l bases reverse-complemented), plus one shorter pair for the remainder (:249-283).
//
// The reference reads its input through btllib::SeqReader (btllib >= 1.4.3, README.md:57; not
// vendored under /root/reference).  What this tool relies on from it: record.id = the name up to
// the first whitespace, record.num = 0-based record index, record.seq / record.qual unchanged, and
// btllib::reverse_complement's table (ACGTU + IUPAC codes, case preserved).  Those are pinned on the
// reference's own golden (Examples/arks-long_test-demo: test_reads.fa.gz -> output/
// test_reads.cut250.fq.gz, byte for byte; tests/test_long_to_linked_pe.py).  Input is parsed with the
// record grammar of seq_reader.h (FASTA/FASTQ, multi-line, gz or plain).
//
// Option handling follows the reference's getopt table including its quirks: `-t` has no `break`
// (:118-120), so its argument also becomes the -f file name; `-v` and `-o` are rejected.
#include "long_cut.h"
#include "seq_reader.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <getopt.h>
#include <iostream>
#include <string>
#include <vector>

namespace {

const std::string PROGNAME = "long-to-linked-pe";
const std::string VERSION = "v1.0";
const size_t MAX_THREADS = 6;

void print_error_msg(const std::string& msg)
{
	std::cerr << PROGNAME << ' ' << VERSION << ": " << msg << std::endl;
}

void print_usage()
{
	std::cerr << "Usage: Split long reads into paired-end pseudo-linked reads." << PROGNAME
	          << " -l L -g G [--fasta -s -d -p P -c C -m M -t T -f FILE -b B --bx / (--bx-only)]  READS "
	             "\n\n"
	             "  -l L        Use L as simulated read length size.\n"
	             "  -g G        Use G as Genome size (bp) for calculating tigmint-long span "
	             "parameter as an integer or in scientific notation (e.g. '3e9').\n"
	             "  --fasta     Output in fasta format.\n"
	             "  -f FILE     Write estimated parameter to FILE. [tigmint-long.params.tsv]\n"
	             "  -s          Calculate span parameter for tigmint-long automatically.\n"
	             "  -c C        Use 'C * sequence coverage' to estimate span parameter. [0.25]\n"
	             "  -d          Calculate dist parameter for tigmint-long automatically.\n"
	             "  -p P        Use P percentile to estimate dist parameter. [50].\n"
	             "  -m M        M minimum read length for a read to be considered a molecule. [2000].\n"
	             "  -t T        Use T number of threads (max 6) per input file. [6]\n"
	             "  -b B        Write barcode multiplicity to B. [barcode_multiplicity.tsv].\n"
	             "  --bx        Compute barcode multiplicity of simulated linked reads output.\n"
	             "  --bx-only  Compute barcode multiplicity of simulated linked reads output only.\n"
	             "  -v          Show verbose output.\n"
	             "  --help      Display this help and exit.\n"
	             "  --version   Display version and exit.\n"
	             "  READS       Space separated list of long reads FASTA/Q files to be cut."
	          << std::endl;
}

// stdout through one large buffer (the reference flushes per record; the bytes are the same)
struct Out
{
	std::string buf;
	Out() { buf.reserve(1u << 22); }
	void flush()
	{
		if (!buf.empty()) {
			fwrite(buf.data(), 1, buf.size(), stdout);
			buf.clear();
		}
	}
	void maybe_flush()
	{
		if (buf.size() > (1u << 21))
			flush();
	}
	~Out()
	{
		flush();
		fflush(stdout);
	}
};

} // namespace

int main(int argc, char* argv[])
{
	int c;
	int optindex = 0;
	static int help = 0, version = 0;
	bool auto_span = false, auto_dist = false;
	size_t l = 0, g = 0, t = 6, m = 2000;
	bool g_set = false;
	bool l_set = false;
	double cov_to_span = 0.25;
	double dist_read_perc = 50;
	size_t dist_lower_bound = 1000;
	std::vector<size_t> read_lengths;
	size_t total_bases = 0;
	static int with_fasta = 0, with_bx_multiplicity = 0, with_bx_multiplicity_only = 0;
	std::string configFile("tigmint-long.params.tsv");
	std::string bxMultiplicityFile("barcode_multiplicity.tsv");
	bool failed = false;
	static const struct option longopts[] = { { "bx", no_argument, &with_bx_multiplicity, 1 },
		                                      { "bx-only", no_argument, &with_bx_multiplicity_only, 1 },
		                                      { "fasta", no_argument, &with_fasta, 1 },
		                                      { "help", no_argument, &help, 1 },
		                                      { "version", no_argument, &version, 1 },
		                                      { nullptr, 0, nullptr, 0 } };
	while ((c = getopt_long(argc, argv, "l:g:o:c:p:sdf:t:b:m:", longopts, &optindex)) != -1) {
		switch (c) {
		case 0: break;
		case 'l':
			l_set = true;
			l = std::stoul(optarg);
			break;
		case 'm': m = std::stoul(optarg); break;
		case 'g':
			g_set = true;
			g = (size_t)std::stod(optarg);
			break;
		case 'p': dist_read_perc = std::stod(optarg); break;
		case 'c': cov_to_span = std::stod(optarg); break;
		case 't':
			t = std::stoul(optarg);
			/* falls through, as in the reference (:118-120) */
			/* fall through */
		case 'f': configFile = optarg; break;
		case 'b': bxMultiplicityFile = optarg; break;
		case 's': auto_span = true; break;
		case 'd': auto_dist = true; break;
		default: std::exit(EXIT_FAILURE);
		}
	}

	std::vector<std::string> infiles(&argv[optind], &argv[argc]);
	if (argc < 2) {
		print_usage();
		std::exit(EXIT_FAILURE);
	}
	if (help != 0) {
		print_usage();
		std::exit(EXIT_SUCCESS);
	} else if (version != 0) {
		std::cerr << PROGNAME << ' ' << VERSION << std::endl;
		std::exit(EXIT_SUCCESS);
	}
	if (!l_set) {
		print_error_msg("missing option -- 'l'");
		failed = true;
	} else if (l == 0) {
		print_error_msg("option has incorrect value -- 'l'");
		failed = true;
	}
	if (!g_set && auto_span) {
		print_error_msg("missing option -- 'g'");
		failed = true;
	} else if (g == 0 && auto_span) {
		print_error_msg("option has incorrect value -- 'g'");
		failed = true;
	}
	if (infiles.empty()) {
		print_error_msg("missing file operand");
		failed = true;
	}
	if (failed) {
		std::cerr << "Try '" << PROGNAME << " --help' for more information.\n";
		std::exit(EXIT_FAILURE);
	}
	if (t > MAX_THREADS) {
		t = MAX_THREADS;
		std::cerr << (PROGNAME + ' ' + VERSION + ": Using more than " + std::to_string(MAX_THREADS) +
		              " threads does not scale, reverting to " + std::to_string(MAX_THREADS) + ".\n")
		          << std::flush;
	}

	std::ofstream bx_multiplicity_ofs;
	if (with_bx_multiplicity_only || with_bx_multiplicity)
		bx_multiplicity_ofs = std::ofstream(bxMultiplicityFile, std::ofstream::out);

	Out out;
	arks_host::LongCutter cutter;
	cutter.l = l;
	cutter.m = m;
	cutter.fasta = with_fasta != 0;
	for (auto& infile : infiles) {
		arks_host::SeqReader reader(infile, 1u << 22);
		if (!reader.ok()) {
			print_error_msg("cannot open " + infile);
			std::exit(EXIT_FAILURE);
		}
		arks_host::SeqRecord record;
		size_t num = 0; // 0-based record index within this file (btllib's record.num)
		for (; reader.read(record) >= 0; ++num) {
			const size_t step = l * 2;
			const std::string& seq = record.seq;
			const size_t seq_size = seq.size();
			if (with_bx_multiplicity_only || with_bx_multiplicity) {
				if (step > seq_size || m > seq_size)
					continue;
				if (seq_size % step != 0)
					bx_multiplicity_ofs << num + 1 << "\t" << (seq_size / step + 1) * 2 << std::endl;
				else
					bx_multiplicity_ofs << num + 1 << "\t" << seq_size / l << std::endl;
			}
			if (with_bx_multiplicity_only)
				continue;
			if (auto_dist && seq_size > dist_lower_bound)
				read_lengths.push_back(seq_size);
			if (auto_span)
				total_bases += seq_size;
			if (step > seq_size || m > seq_size)
				continue;

			cutter.append_pairs(out.buf, record, num);
			out.maybe_flush();
		}
	}
	out.flush();

	if (auto_span || auto_dist) {
		std::ofstream ofs(configFile, std::ofstream::app);
		if (auto_span)
			ofs << "span\t" << (size_t)(total_bases / g * cov_to_span) << "\n";
		if (auto_dist) {
			if (read_lengths.size() == 0) {
				std::cerr << "long-to-linked-pe: unable to estimate dist parameter due to no valid "
				             "lengths"
				          << std::endl;
			} else {
				size_t dist_estimate;
				std::sort(read_lengths.begin(), read_lengths.end());
				double index = (dist_read_perc / 100) * read_lengths.size();
				size_t size_t_index = (size_t)floor(index);
				if (floor(index) == index)
					dist_estimate = (read_lengths[size_t_index - 1] + read_lengths[size_t_index]) / 2;
				else
					dist_estimate = read_lengths[size_t_index];
				ofs << "read_p" << dist_read_perc << "\t" << dist_estimate << "\n";
			}
		}
		ofs.close();
	}
	return 0;
}
