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

void complain(const std::string& what)
{
	std::cerr << PROGNAME << ' ' << VERSION << ": " << what << std::endl;
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

// everything the command line says
struct Options
{
	size_t read_len = 0;      // -l
	bool have_read_len = false;
	size_t genome_size = 0;   // -g (integer or scientific notation)
	bool have_genome_size = false;
	size_t threads = 6;       // -t (accepted; this tool is not bound by the cutting)
	size_t min_long = 2000;   // -m
	double span_factor = 0.25; // -c
	double dist_percentile = 50; // -p
	bool want_span = false, want_dist = false; // -s, -d
	std::string params_file = "tigmint-long.params.tsv"; // -f
	std::string mult_file = "barcode_multiplicity.tsv";  // -b
	int fasta = 0, bx = 0, bx_only = 0, help = 0, version = 0;
	std::vector<std::string> inputs;
};

// getopt table of the reference (src/long-to-linked-pe.cpp:86-139); an option it does not handle ends the program
Options parse_command_line(int argc, char* argv[])
{
	static Options o; // (getopt_long stores the flags through pointers)
	const struct option table[] = { { "bx", no_argument, &o.bx, 1 },
		                            { "bx-only", no_argument, &o.bx_only, 1 },
		                            { "fasta", no_argument, &o.fasta, 1 },
		                            { "help", no_argument, &o.help, 1 },
		                            { "version", no_argument, &o.version, 1 },
		                            { nullptr, 0, nullptr, 0 } };
	int idx = 0;
	for (int c; (c = getopt_long(argc, argv, "l:g:o:c:p:sdf:t:b:m:", table, &idx)) != -1;) {
		if (c == 0)
			continue;
		if (c == 'l') {
			o.read_len = std::stoul(optarg);
			o.have_read_len = true;
		} else if (c == 'g') {
			o.genome_size = (size_t)std::stod(optarg);
			o.have_genome_size = true;
		} else if (c == 'm') {
			o.min_long = std::stoul(optarg);
		} else if (c == 'p') {
			o.dist_percentile = std::stod(optarg);
		} else if (c == 'c') {
			o.span_factor = std::stod(optarg);
		} else if (c == 't' || c == 'f') {
			// the reference's `case 't'` has no break (:118-120): the thread count also becomes the -f file name
			if (c == 't')
				o.threads = std::stoul(optarg);
			o.params_file = optarg;
		} else if (c == 'b') {
			o.mult_file = optarg;
		} else if (c == 's') {
			o.want_span = true;
		} else if (c == 'd') {
			o.want_dist = true;
		} else {
			std::exit(EXIT_FAILURE); // unknown options, and -o, which is in the option string without a case
		}
	}
	o.inputs.assign(argv + optind, argv + argc);
	return o;
}

// the reference's checks, in its order and with its messages (:141-182)
void validate_or_exit(int argc, Options& o)
{
	if (argc < 2) {
		print_usage();
		std::exit(EXIT_FAILURE);
	}
	if (o.help) {
		print_usage();
		std::exit(EXIT_SUCCESS);
	}
	if (o.version) {
		std::cerr << PROGNAME << ' ' << VERSION << std::endl;
		std::exit(EXIT_SUCCESS);
	}
	int problems = 0;
	if (!o.have_read_len)
		complain("missing option -- 'l'"), ++problems;
	else if (o.read_len == 0)
		complain("option has incorrect value -- 'l'"), ++problems;
	if (o.want_span && !o.have_genome_size)
		complain("missing option -- 'g'"), ++problems;
	else if (o.want_span && o.genome_size == 0)
		complain("option has incorrect value -- 'g'"), ++problems;
	if (o.inputs.empty())
		complain("missing file operand"), ++problems;
	if (problems) {
		std::cerr << "Try '" << PROGNAME << " --help' for more information.\n";
		std::exit(EXIT_FAILURE);
	}
	if (o.threads > MAX_THREADS) {
		o.threads = MAX_THREADS;
		std::cerr << (PROGNAME + ' ' + VERSION + ": Using more than " + std::to_string(MAX_THREADS) +
		              " threads does not scale, reverting to " + std::to_string(MAX_THREADS) + ".\n")
		          << std::flush;
	}
}

// what -s / -d append to the tigmint-long parameter file (:294-322)
struct LengthStats
{
	std::vector<size_t> lengths; // of the reads longer than 1000 bases (-d)
	size_t bases = 0;            // of all reads (-s)

	void write(const Options& o) const
	{
		std::ofstream f(o.params_file, std::ofstream::app);
		if (o.want_span)
			f << "span\t" << (size_t)(bases / o.genome_size * o.span_factor) << "\n";
		if (!o.want_dist)
			return;
		if (lengths.empty()) {
			std::cerr << "long-to-linked-pe: unable to estimate dist parameter due to no valid "
			             "lengths"
			          << std::endl;
			return;
		}
		std::vector<size_t> sorted(lengths);
		std::sort(sorted.begin(), sorted.end());
		// the percentile as the reference takes it: the mean of two neighbours when the rank is whole
		const double rank = (o.dist_percentile / 100) * sorted.size();
		const size_t at = (size_t)floor(rank);
		const size_t estimate = floor(rank) == rank ? (sorted[at - 1] + sorted[at]) / 2 : sorted[at];
		f << "read_p" << o.dist_percentile << "\t" << estimate << "\n";
	}
};

} // namespace

int main(int argc, char* argv[])
{
	Options opt = parse_command_line(argc, argv);
	validate_or_exit(argc, opt);

	const bool count_bx = opt.bx || opt.bx_only;
	std::ofstream mult_out;
	if (count_bx)
		mult_out.open(opt.mult_file, std::ofstream::out);

	arks_host::LongCutter cutter;
	cutter.l = opt.read_len;
	cutter.m = opt.min_long;
	cutter.fasta = opt.fasta != 0;
	const size_t dist_lower_bound = 1000;
	LengthStats stats;
	Out out;
	for (const std::string& path : opt.inputs) {
		arks_host::SeqReader reader(path, 1u << 22);
		if (!reader.ok()) {
			complain("cannot open " + path);
			std::exit(EXIT_FAILURE);
		}
		arks_host::SeqRecord rec;
		// num: 0-based record index within this file (btllib's record.num); the barcode is num + 1
		for (size_t num = 0; reader.read(rec) >= 0; ++num) {
			const size_t len = rec.seq.size();
			const bool cut_it = cutter.accepts(len);
			if (count_bx) {
				if (!cut_it)
					continue;
				mult_out << num + 1 << "\t" << cutter.multiplicity(len) << std::endl;
				if (opt.bx_only)
					continue;
			}
			if (opt.want_dist && len > dist_lower_bound)
				stats.lengths.push_back(len);
			if (opt.want_span)
				stats.bases += len;
			if (!cut_it)
				continue;
			cutter.append_pairs(out.buf, rec, num);
			out.maybe_flush();
		}
	}
	out.flush();
	if (opt.want_span || opt.want_dist)
		stats.write(opt);
	return 0;
}
