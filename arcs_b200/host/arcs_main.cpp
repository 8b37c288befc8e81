// arcs_main.cpp -- `arcs --arks` drop-in: same flags, defaults, input formats and output files as
// bcgsc/arcs 1.2.8 (Arcs/Arcs.cpp main/runArcs, :1810-2184), with the ARKS hot path
// (contig-end k-mer index, read-pair lookup + vote, barcode tally, pair links) running on the
// GPU through the C ABI of libarks_b200.so.  Host side = parsing, barcode interning, graph +
// writers.  No CPU fallback for the hot path.
//
// Differences from the reference, all outside the result files:
//  * reads are parsed ONCE: barcode multiplicities are counted while mapping; because the
//    reference's only use of the multiplicity before pairContigs is "is this barcode known"
//    (Arcs.cpp:1257-1267), mapping first and filtering later gives the same imap.  If -m is
//    given as an inverted range (min >= max), where the reference's `goodmult` test can fail,
//    the reads are counted in a separate first pass exactly as the reference does.
//  * -t is accepted and ignored (the GPU does the mapping); --gpus N (or ARKS_GPUS) shards
//    read pairs by barcode across N GPUs.
//  * -D (distance estimation, Arcs/DistanceEst.h) runs on the host over the IndexMap rows exported by
//    the GPU.  Two of its results depend on the iteration order of std::unordered_map containers in the
//    reference (which intra-contig sample survives when two contigs have the same barcode Jaccard index,
//    and the line order of --samples_tsv); the same containers are filled in the same order here
//    (barcodes in the order their first pair was stored, which the GPU reports per pair under -D), so
//    the output equals the reference's at -t 1 built with the same libstdc++.
//  * the alignment mode (no --arks: SAM text in, Arcs.cpp:572-771) tallies read pairs on the host and hands
//    the IndexMap rows to the same GPU pair-link kernel.
#include "../../include/arks_b200.h"
#include "fasta_fast.h"
#include "ingest.h"
#include "long_cut.h"
#include "seq_reader.h"

#include <algorithm>
#include <cassert>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <getopt.h>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <unistd.h>
#include <unordered_map>
#include <vector>

#define PROGRAM "arcs"
#define PACKAGE_VERSION "1.2.8-b200"

using arks_host::Barcodes;
using arks_host::count_barcode;
using arks_host::extract_bx;
using arks_host::SeqReader;
using arks_host::SeqRecord;
using arks_host::strip_read_num;

namespace {

// ---- parameters: ARCS::ArcsParams (Arcs/Arcs.h:43-103), same defaults -----------------------
struct Params
{
	std::string file, fofName, base_name, dist_graph_name, tsv_name, barcode_counts_name, multfile;
	std::string dist_samples_tsv, dist_tsv;
	int seq_id = 98, min_reads = 5, min_links = 0, min_size = 500, min_mult = 50, max_mult = 10000, max_degree = 0,
	    end_length = 30000, verbose = 0, k_value = 30;
	unsigned gap = 100, threads = 1, dist_bin_size = 20;
	float error_percent = 0.05f;
	double j_index = 0.55;
	bool arks = false, dist_est = false, output_pair = false, dist_upper = false;
	int gpus = 1;
	bool two_pass = false;
	size_t cut = 0, cut_min = 2000; // --cut / --cut_min: the inputs are long reads, cut on the fly (long_cut.h)
};
Params params;

const char USAGE[] =
    "Usage: " PROGRAM " --arks -f CONTIGS.fa [OPTION]... READS.fq[.gz]...\n"
    "B200-native ARKS (k-mer) mode of ARCS.  Options (as in arcs " PACKAGE_VERSION "):\n"
    "   -f, --file=FILE       FASTA file of contig sequences to scaffold (required)\n"
    "   -a, --fofName=FILE    text file listing input read files\n"
    "   -u, --multfile=FILE   tsv or csv file listing barcode multiplicities\n"
    "   -c, --min_reads=N     min aligned read pairs per barcode mapping [5]\n"
    "   -k, --k_value=N       size of a k-mer [30] (4..64)\n"
    "   -j, --j_index=N       min fraction of read kmers matching a contigId [0.55]\n"
    "   -t, --threads=N       accepted for compatibility; the GPU does the mapping\n"
    "   -l, --min_links=N     min shared barcodes between contigs [0]\n"
    "   -z, --min_size=N      min contig length [500]\n"
    "   -b, --base_name=STR   output file prefix\n"
    "   -g, --graph=FILE      write the ABySS dist.gv to FILE\n"
    "       --gap=N           fixed gap size for ABySS dist.gv file [100]\n"
    "       --tsv=FILE        write graph in TSV format to FILE\n"
    "       --barcode-counts=FILE       write number of reads per barcode to FILE\n"
    "   -m, --index_multiplicity=RANGE  barcode multiplicity range [50-10000]\n"
    "   -d, --max_degree=N    max node degree in scaffold graph [0]\n"
    "   -e, --end_length=N    contig head/tail length for masking alignments [30000]\n"
    "   -r, --error_percent=N p-value for head/tail assignment and link orientation [0.05]\n"
    "   -P, --pair            output scaffolds pairing TSV\n"
    "   -v, --run_verbose     verbose logging\n"
    "       --gpus=N          number of GPUs to shard read pairs over [1]\n"
    "       --cut=L           the read files hold long reads: cut them into pseudo-linked read pairs of\n"
    "                         L bases while reading (what `long-to-linked-pe -l L` would pipe in)\n"
    "       --cut_min=M       with --cut: minimum length of a long read [2000]\n"
    "       --arks            k-mer method (required)\n";

enum
{
	OPT_HELP = 1000,
	OPT_VERSION,
	OPT_BX,
	OPT_GAP,
	OPT_TSV,
	OPT_BARCODE_COUNTS,
	OPT_SAMPLES_TSV,
	OPT_DIST_TSV,
	OPT_NO_DIST_EST,
	OPT_DIST_MEDIAN,
	OPT_DIST_UPPER,
	OPT_ARKS_METHOD,
	OPT_GPUS,
	OPT_TWO_PASS,
	OPT_CUT,
	OPT_CUT_MIN
};

const char shortopts[] = "f:a:B:s:c:Dl:z:b:g:m:d:e:r:vt:u:j:k:P";
const struct option longopts[] = { { "file", required_argument, NULL, 'f' },
	                               { "fofName", required_argument, NULL, 'a' },
	                               { "bin_size", required_argument, NULL, 'B' },
	                               { "bx", no_argument, NULL, OPT_BX },
	                               { "samples_tsv", required_argument, NULL, OPT_SAMPLES_TSV },
	                               { "dist_tsv", required_argument, NULL, OPT_DIST_TSV },
	                               { "seq_id", required_argument, NULL, 's' },
	                               { "min_reads", required_argument, NULL, 'c' },
	                               { "dist_est", no_argument, NULL, 'D' },
	                               { "no_dist_est", no_argument, NULL, OPT_NO_DIST_EST },
	                               { "dist_median", no_argument, NULL, OPT_DIST_MEDIAN },
	                               { "dist_upper", no_argument, NULL, OPT_DIST_UPPER },
	                               { "min_links", required_argument, NULL, 'l' },
	                               { "min_size", required_argument, NULL, 'z' },
	                               { "base_name", required_argument, NULL, 'b' },
	                               { "graph", required_argument, NULL, 'g' },
	                               { "tsv", required_argument, NULL, OPT_TSV },
	                               { "barcode-counts", required_argument, NULL, OPT_BARCODE_COUNTS },
	                               { "gap", required_argument, NULL, OPT_GAP },
	                               { "index_multiplicity", required_argument, NULL, 'm' },
	                               { "max_degree", required_argument, NULL, 'd' },
	                               { "end_length", required_argument, NULL, 'e' },
	                               { "error_percent", required_argument, NULL, 'r' },
	                               { "run_verbose", required_argument, NULL, 'v' },
	                               { "version", no_argument, NULL, OPT_VERSION },
	                               { "help", no_argument, NULL, OPT_HELP },
	                               { "threads", required_argument, NULL, 't' },
	                               { "multfile", required_argument, NULL, 'u' },
	                               { "k_value", required_argument, NULL, 'k' },
	                               { "j_index", required_argument, NULL, 'j' },
	                               { "arks", no_argument, NULL, OPT_ARKS_METHOD },
	                               { "pair", no_argument, NULL, 'P' },
	                               { "gpus", required_argument, NULL, OPT_GPUS },
	                               { "two-pass", no_argument, NULL, OPT_TWO_PASS },
	                               { "cut", required_argument, NULL, OPT_CUT },
	                               { "cut_min", required_argument, NULL, OPT_CUT_MIN },
	                               { NULL, 0, NULL, 0 } };

// A read file as a byte stream: the file itself, or -- with --cut -- the pseudo-linked reads cut from the
// long reads in it.  Exits like the reference when the file cannot be opened (Arcs.cpp:1166-1170).
std::unique_ptr<arks_host::ByteSource> open_reads(const std::string& f)
{
	std::unique_ptr<arks_host::ByteSource> src;
	if (params.cut) {
		std::unique_ptr<arks_host::LongCutSource> cut(new arks_host::LongCutSource(f, params.cut, params.cut_min));
		if (cut->ok())
			src = std::move(cut);
	} else {
		src = arks_host::open_source(f);
	}
	if (!src) {
		std::cerr << "File " << f << " cannot be opened." << std::endl;
		exit(1);
	}
	return src;
}

[[noreturn]] void die(const std::string& msg)
{
	std::cerr << PROGRAM ": " << msg << std::endl;
	exit(EXIT_FAILURE);
}

void assert_readable(const std::string& path)
{
	if (access(path.c_str(), R_OK) == -1) {
		std::cerr << "error: `" << path << "': " << strerror(errno) << std::endl;
		exit(EXIT_FAILURE);
	}
}

const char* maybeNA(const std::string& s)
{
	return s.empty() ? "NA" : s.c_str();
}

std::string stamp()
{
	std::time_t t;
	time(&t);
	return ctime(&t);
}

double now()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- small restatements of the reference's helpers -------------------------------------------

// readFof (Arcs.cpp:776-794): one file name per whitespace-separated token
std::vector<std::string> read_fof(const std::string& fof)
{
	std::vector<std::string> out;
	if (fof.empty())
		return out;
	std::ifstream in(fof.c_str());
	if (!in)
		die("error: could not open `" + fof + "'");
	std::string tok;
	while (in >> tok)
		out.push_back(tok);
	return out;
}

// checkSameFormat (Arcs.cpp:336-361)
bool check_same_format(const std::vector<std::string>& files, bool& all_alignment)
{
	int prev = 0, cur = 0;
	for (const auto& f : files) {
		cur = 0;
		if (f.find(".sam") != std::string::npos || f.find(".bam") != std::string::npos)
			cur = 1;
		if (f.find(".fastq") != std::string::npos || f.find(".fq") != std::string::npos)
			cur = 2;
		if (params.cut && !cur && f.find(".fa") != std::string::npos) // long reads to cut (--cut) also come as FASTA
			cur = 2;
		if (!cur) {
			std::cout << "Unknown type file is observed!" << std::endl;
			return false;
		}
		if (prev && prev != cur)
			return false;
		prev = cur;
	}
	all_alignment = (cur == 1);
	return true;
}

// normalEstimation / checkSignificance (Arcs.cpp:833-839,1459-1467), same types as the reference
float normal_estimation(int x, float p, int n)
{
	float mean = n * p;
	float sd = std::sqrt(n * p * (1 - p));
	return 0.5 * (1 + std::erf((x - mean) / (sd * std::sqrt(2))));
}

bool check_significance(int max, int second)
{
	if (max < params.min_links)
		return false;
	float cdf = normal_estimation(max, 0.5, second);
	return (1 - cdf < params.error_percent);
}

// ---- barcodes (struct Barcodes: ingest.h) -----------------------------------------------------------
// createIndexMultMap (Arcs.cpp:392-448)
void load_multfile(const std::string& path, Barcodes& bc)
{
	const bool tsv = path.find(".tsv") != std::string::npos;
	std::ifstream in(path.c_str());
	if (!in) {
		std::cerr << "Could not open " << path << ". --fatal.\n";
		exit(EXIT_FAILURE);
	}
	std::string line;
	size_t n = 0;
	while (getline(in, line)) {
		std::string barcode, ms;
		if (tsv) {
			std::stringstream ss(line);
			ss >> barcode >> ms;
		} else {
			std::istringstream iss(line);
			getline(iss, barcode, ',');
			iss >> ms;
		}
		n++;
		int m = std::stoi(ms);
		if (!barcode.empty()) {
			uint32_t i = bc.intern(barcode);
			bc.mult[i] = m;
			bc.counted[i] = 1;
		} else {
			std::cout << "Please check your multiplicity file." << std::endl;
		}
	}
	if (params.verbose)
		std::cout << "Saw " << n << "  distinct barcodes." << std::endl;
}

// ---- contigs -----------------------------------------------------------------------------------
struct Contigs
{
	std::vector<std::string> name;           // kept contigs (len >= -z), file order; conreci 2i+1 / 2i+2
	std::vector<int> length;
	std::unordered_map<std::string, int> to_length; // ARCS::ContigToLength (iteration order matters for .dist.gv)
	std::vector<uint32_t> first_of_name;     // index of the first kept contig with the same name
	size_t total = 0, skipped = 0;
};

// ---- GPU handles -------------------------------------------------------------------------------
struct Gpu
{
	arks_handle* h = nullptr;
	// two pinned batches
	struct Batch
	{
		char* bases = nullptr;
		uint32_t* off = nullptr;
		uint32_t* bc = nullptr;
		uint64_t n_bases = 0;
		uint32_t n_pairs = 0;
		int32_t* out = nullptr;    // -D only: the contig end each pair was stored under (0: not stored)
		std::vector<uint64_t> seq; // -D only: position of each pair in the input
	} batch[2];
	int cur = 0;
};

// -D: position in the input of the first STORED pair of each barcode = the order in which the reference
// inserts barcodes into its IndexMap (Arcs.cpp:1280-1285)
std::vector<uint64_t> g_first_stored;
uint64_t g_pair_seq = 0;

constexpr uint64_t kBatchBases = 256ull << 20;
constexpr uint32_t kBatchPairs = 1u << 20;

// Multi-GPU hosts: the thread that feeds GPU d (and the pinned buffers it touches first) should live on the
// NUMA node the GPU hangs off.  arks_bind_thread pins the calling thread to that node's cores; it does nothing
// on single-node hosts or when the topology cannot be read.
void bind_thread_to_gpu(int d)
{
	if (params.gpus > 1 && !getenv("ARKS_NO_BIND") && !getenv("ARKS_GPUS_SAME_DEVICE"))
		arks_bind_thread(d);
}

void ck(arks_handle* h, int rc, const char* what)
{
	if (rc != ARKS_OK) {
		std::cerr << PROGRAM ": GPU error in " << what << " (" << rc << "): " << arks_last_error(h) << std::endl;
		exit(EXIT_FAILURE);
	}
}

void flush_batch(Gpu& g)
{
	Gpu::Batch& b = g.batch[g.cur];
	if (b.n_pairs == 0)
		return;
	b.off[2 * b.n_pairs] = (uint32_t)b.n_bases;
	ck(g.h, arks_map_pairs(g.h, b.bases, b.off, b.bc, b.n_pairs, params.j_index, b.out), "arks_map_pairs");
	if (b.out) { // the call is synchronous when it returns per-pair results
		for (uint32_t i = 0; i < b.n_pairs; ++i)
			if (b.out[i] != 0) {
				const uint32_t id = b.bc[i];
				if (id >= g_first_stored.size())
					g_first_stored.resize((size_t)id + 1 + g_first_stored.size() / 2, UINT64_MAX);
				g_first_stored[id] = std::min(g_first_stored[id], b.seq[i]);
			}
		b.seq.clear();
	}
	g.cur ^= 1;
	g.batch[g.cur].n_pairs = 0;
	g.batch[g.cur].n_bases = 0;
}

// the two large pinned batches of the record-by-record path, allocated on first use (input that goes
// through the block path never needs them, and pinning 0.5 GB takes a noticeable fraction of a second)
void ensure_batches(Gpu& g)
{
	if (g.batch[0].bases)
		return;
	for (auto& b : g.batch) {
		void* p;
		if (arks_host_alloc(&p, kBatchBases + 64) != ARKS_OK)
			die("error: cannot allocate pinned host memory");
		b.bases = (char*)p;
		if (arks_host_alloc(&p, (2ull * kBatchPairs + 1) * 4) != ARKS_OK)
			die("error: cannot allocate pinned host memory");
		b.off = (uint32_t*)p;
		if (arks_host_alloc(&p, kBatchPairs * 4ull) != ARKS_OK)
			die("error: cannot allocate pinned host memory");
		b.bc = (uint32_t*)p;
		if (params.dist_est) {
			if (arks_host_alloc(&p, kBatchPairs * 4ull) != ARKS_OK)
				die("error: cannot allocate pinned host memory");
			b.out = (int32_t*)p;
		}
	}
}

void add_pair(Gpu& g, const char* s1, size_t l1, const char* s2, size_t l2, uint32_t barcode)
{
	if (l1 + l2 > kBatchBases / 2)
		die("error: read pair longer than the batch buffer");
	ensure_batches(g);
	Gpu::Batch* b = &g.batch[g.cur];
	if (b->n_pairs >= kBatchPairs || b->n_bases + l1 + l2 > kBatchBases) {
		flush_batch(g);
		b = &g.batch[g.cur];
	}
	b->off[2 * b->n_pairs] = (uint32_t)b->n_bases;
	memcpy(b->bases + b->n_bases, s1, l1);
	b->n_bases += l1;
	b->off[2 * b->n_pairs + 1] = (uint32_t)b->n_bases;
	memcpy(b->bases + b->n_bases, s2, l2);
	b->n_bases += l2;
	b->bc[b->n_pairs] = barcode;
	if (b->out)
		b->seq.push_back(g_pair_seq);
	g_pair_seq++;
	b->n_pairs++;
}

// ---- graph (createGraph / removeDegreeNodes / write_graphviz without Boost) --------------------
struct Edge
{
	int u, v, orientation, weight;
	// EdgeProperties' defaults (Arcs/Arcs.h:166-182); set by -D
	int minDist = INT_MIN, dist = INT_MAX, maxDist = INT_MAX;
	float jaccard = -1.0f;
};
struct Graph
{
	std::vector<std::string> vid; // vertex -> contig name
	std::vector<Edge> edges;      // in pmap order
};

// ---- ARCS alignment mode: readBAM (Arcs/Arcs.cpp:572-771) -----------------------------------------
// SAM text (the pipeline feeds `samtools view -h`), read pairs on consecutive lines.  A pair whose two
// alignments pass the flag / MAPQ / identity tests and hit the same contig is tallied under its barcode
// at the head or the tail of that contig by the mean of the two positions -- when the NEXT read name
// arrives (so the last pair of a file is never tallied, as upstream).  The tallies become IndexMap rows
// for the GPU pair-link kernel (arks_imap_add).
struct AlignRows
{
	std::unordered_map<uint64_t, std::pair<uint32_t, uint32_t>> ht; // barcode << 32 | contig -> head, tail
	std::vector<uint64_t> first_add;                                 // per barcode: order of its first tally
	uint64_t n_adds = 0;
};

uint32_t intern_contig(Contigs& ct, std::unordered_map<std::string, uint32_t>& contig_id, const std::string& name, int length, bool overwrite)
{
	auto it = contig_id.find(name);
	if (it != contig_id.end()) {
		if (overwrite) {
			ct.length[it->second] = length;
			ct.to_length[name] = length;
		}
		return it->second;
	}
	const uint32_t i = (uint32_t)ct.name.size();
	contig_id.emplace(name, i);
	ct.name.push_back(name);
	ct.length.push_back(length);
	ct.first_of_name.push_back(i);
	ct.to_length[name] = length;
	return i;
}

// checkFlag (Arcs.cpp:204-211): PAIRED,PROPER_PAIR with exactly one of REVERSE / MREVERSE
bool accepted_flag(int flag)
{
	flag &= ~0xc0;
	return flag == 19 || flag == 35;
}

// calcSequenceIdentity (Arcs.cpp:275-315): aligned query length (M,=,X,I) minus NM, over the read length, in percent
double sequence_identity(const std::string& line, const std::string& cigar, const std::string& seq)
{
	int qalen = 0;
	long num = 0;
	bool have_digits = false, broken = false;
	for (char c : cigar) {
		if (isdigit((unsigned char)c)) {
			num = num * 10 + (c - '0');
			have_digits = true;
		} else {
			if (c == 'M' || c == '=' || c == 'X' || c == 'I') {
				if (!have_digits)
					broken = true; // upstream's stream extraction fails here and stays failed
				if (!broken)
					qalen += (int)num;
			}
			num = 0;
			have_digits = false;
		}
	}
	int edit_dist = 0;
	const size_t found = line.find("NM:i:");
	if (found != std::string::npos)
		edit_dist = (int)std::strtol(&line[found + 5], 0, 10);
	double si = 0;
	if (qalen != 0) {
		const double mins = qalen - edit_dist;
		si = mins / seq.length() * 100;
	}
	return si;
}

// parseBXTag (Common/SAM.h:9-33): first "BX:Z:" up to the next blank, tab, CR or LF
std::string sam_bx_tag(const std::string& s)
{
	size_t start = s.find("BX:Z:");
	if (start == std::string::npos)
		return std::string();
	start += 5;
	size_t end = s.find_first_of(" \t\r\n", start);
	if (end == std::string::npos)
		end = s.length();
	return s.substr(start, end - start);
}

void read_alignments(const std::string& path, Barcodes& bc, Contigs& ct, std::unordered_map<std::string, uint32_t>& contig_id, AlignRows& rows)
{
	std::ifstream in(path.c_str());
	if (!in.good()) {
		std::cerr << "error: `" << path << "': " << strerror(errno) << std::endl;
		exit(EXIT_FAILURE);
	}
	if (in.peek() == EOF) {
		std::cerr << "error: alignments file is empty: " << path << '\n';
		exit(EXIT_FAILURE);
	}
	std::string prevRN, readyIndex, prevRef, readyRef;
	int prevSI = 0, prevFlag = 0, prevMapq = 0, prevPos = -1, readyPos = -1;
	int nth = 1; // which alignment of the current read name this line is
	size_t linecount = 0, countUnpaired = 0;
	const bool add_sq_lengths = ct.to_length.empty();
	std::string line;
	while (getline(in, line)) {
		if (line.empty())
			continue;
		if (line[0] == '@') {
			if (line.compare(0, 4, "@SQ\t") == 0) {
				// "@SQ\tSN:<name>\tLN:<size>" (upstream reads it with `expect` manipulators and exits on anything else)
				std::string name;
				size_t size = 0;
				bool ok = line.compare(0, 7, "@SQ\tSN:") == 0;
				size_t p = 7;
				if (ok) {
					while (p < line.size() && !isspace((unsigned char)line[p]))
						name.push_back(line[p++]);
					ok = !name.empty() && line.compare(p, 4, "\tLN:") == 0;
					p += 4;
				}
				if (ok) {
					size_t q = p;
					while (q < line.size() && isdigit((unsigned char)line[q]))
						size = size * 10 + (size_t)(line[q++] - '0');
					ok = q > p;
				}
				if (!ok) {
					std::cerr << "error: parsing SAM header: " << line << '\n';
					exit(EXIT_FAILURE);
				}
				if (add_sq_lengths) {
					if (!contig_id.count(name))
						intern_contig(ct, contig_id, name, (int)size, false);
				} else {
					auto it = ct.to_length.find(name);
					if (it == ct.to_length.end()) {
						std::cerr << "error: unexpected sequence: " << name << " of size " << size;
						exit(EXIT_FAILURE);
					} else if (it->second != (int)size) {
						std::cerr << "error: mismatched sequence lengths: sequence " << name << ": " << it->second << " != " << size;
						exit(EXIT_FAILURE);
					}
				}
			}
			continue;
		}
		linecount++;
		std::stringstream ss(line);
		std::string readName, scafName, cigar, rnext, seq, qual, tags;
		int flag = 0, pos = 0, mapq = 0, pnext = 0, tlen = 0;
		ss >> readName >> flag >> scafName >> pos >> mapq >> cigar >> rnext >> pnext >> tlen >> seq >> qual >> std::ws;
		getline(ss, tags);
		// barcode: BX tag, else the suffix of the read name after the last '_' if it is all ACGT
		std::string index = sam_bx_tag(tags);
		if (index.empty()) {
			const size_t found = readName.rfind("_");
			if (found != std::string::npos) {
				index = readName.substr(found + 1);
				if (index.find_first_not_of("ACGTacgt") != std::string::npos)
					index.clear();
			}
		}
		// multiplicity counts primary, non-supplementary alignments
		if (!index.empty() && !((flag & 0x800) || (flag & 0x100))) {
			const uint32_t b = bc.intern(index);
			bc.mult[b]++;
			bc.counted[b] = 1;
		}
		const int si = (int)sequence_identity(line, cigar, seq);
		if (nth == 2 && readName != prevRN) {
			if (countUnpaired == 0)
				std::cerr << "Warning: Skipping an unpaired read. Read pairs should be consecutive in the SAM/BAM file.\n"
				             "  Prev read: "
				          << prevRN << "\n  Curr read: " << readName << std::endl;
			++countUnpaired;
			if (countUnpaired % 1000000 == 0)
				std::cerr << "Warning: Skipped " << countUnpaired << " unpaired reads." << std::endl;
			nth = 1;
		}
		if (nth >= 3)
			nth = 1;
		if (nth == 1) {
			if (readName != prevRN) {
				prevRN = readName;
				prevSI = si;
				prevFlag = flag;
				prevMapq = mapq;
				prevRef = scafName;
				prevPos = pos;
				// the previous read name had exactly two alignments: tally its pair now
				if (!readyIndex.empty() && !readyRef.empty() && readyRef != "*" && readyPos != -1) {
					// (upstream's operator[] creates a zero-length entry for a contig it has never seen)
					const uint32_t c = intern_contig(ct, contig_id, readyRef, 0, false);
					const int size = ct.to_length[readyRef];
					if (size >= params.min_size) {
						int cutOff = params.end_length;
						if (cutOff == 0 || size <= cutOff * 2)
							cutOff = size / 2;
						const bool head = readyPos <= cutOff, tail = !head && readyPos > size - cutOff;
						if (head || tail) {
							const uint32_t b = bc.intern(readyIndex);
							auto& ht = rows.ht[((uint64_t)b << 32) | c];
							(head ? ht.first : ht.second)++;
							if (b >= rows.first_add.size())
								rows.first_add.resize((size_t)b + 1 + rows.first_add.size() / 2, UINT64_MAX);
							rows.first_add[b] = std::min(rows.first_add[b], rows.n_adds);
							rows.n_adds++;
						}
					}
					readyIndex.clear();
					readyRef.clear();
					readyPos = -1;
				}
			} else { // a third alignment under the same name: the name is dropped
				nth = 0;
				readyIndex.clear();
				readyRef.clear();
				readyPos = -1;
			}
		} else if (nth == 2) {
			if (!seq.empty() && accepted_flag(flag) && accepted_flag(prevFlag) && mapq != 0 && prevMapq != 0 && si >= params.seq_id &&
			    prevSI >= params.seq_id) {
				if (prevRef == scafName && scafName != "*" && !scafName.empty() && !index.empty()) {
					readyIndex = index;
					readyRef = scafName;
					readyPos = (prevPos + pos) / 2;
				}
			}
		}
		nth++;
		if (params.verbose && linecount % 10000000 == 0)
			std::cout << "On line " << linecount << std::endl;
	}
	if (countUnpaired > 0)
		std::cerr << "Warning: Skipped " << countUnpaired << " unpaired reads. Read pairs should be consecutive in the SAM/BAM file.\n";
}

// ---- -D: distance estimation (Arcs/DistanceEst.h) over the exported IndexMap rows ----------------
struct DistSample // DistanceEst.h:37-53
{
	unsigned distance = UINT_MAX, barcodesHead = 0, barcodesTail = 0, barcodesUnion = 0, barcodesIntersect = 0;
};
struct BarcodeStats // DistanceEst.h:64-78
{
	unsigned barcodes1 = 0, barcodes2 = 0, barcodesUnion = 0, barcodesIntersect = 0;
};

struct DistInput
{
	const std::vector<uint32_t>&im_bc, &im_ct, &im_h, &im_t; // IndexMap rows: barcode, contig, head count, tail count
	const std::vector<std::string>& bc_name;
	const std::vector<int32_t>& bc_mult;
	const std::vector<std::string>& ct_name;
	const std::unordered_map<std::string, int>& to_length;
	const std::vector<uint32_t>& lexrank;
	const std::vector<uint64_t>& first_stored;
};

// quantile (Common/StatUtil.h:9-33), including its weighting (the element BEFORE the boundary gets
// the fractional part)
double quantile(const std::vector<unsigned>& v, double q)
{
	const size_t lastPos = v.size() - 1;
	const size_t beforePos = (size_t)floor(q * lastPos);
	const size_t before = v[beforePos];
	const size_t afterPos = (size_t)ceil(q * lastPos);
	const size_t after = v[afterPos];
	const double weight = (q * lastPos - beforePos) / 1.0;
	return weight * before + (1.0 - weight) * after;
}

void calc_distance_estimates(const DistInput& in, Graph& g)
{
	const size_t n_rows = in.im_bc.size();
	const unsigned two_e = (unsigned)2 * params.end_length;
	// rows grouped by barcode, each barcode's rows in ScafMap order (contig name under std::string '<';
	// per contig the tail entry (name, false) comes before the head entry (name, true))
	std::vector<uint32_t> rows(n_rows);
	for (size_t i = 0; i < n_rows; ++i)
		rows[i] = (uint32_t)i;
	std::sort(rows.begin(), rows.end(), [&](uint32_t a, uint32_t b) {
		return in.im_bc[a] != in.im_bc[b] ? in.im_bc[a] < in.im_bc[b] : in.lexrank[in.im_ct[a]] < in.lexrank[in.im_ct[b]];
	});
	std::unordered_map<uint32_t, std::pair<size_t, size_t>> range_of; // barcode -> [first, last) in rows
	for (size_t i = 0; i < n_rows;) {
		size_t j = i;
		while (j < n_rows && in.im_bc[rows[j]] == in.im_bc[rows[i]])
			++j;
		range_of[in.im_bc[rows[i]]] = std::make_pair(i, j);
		i = j;
	}
	// the reference iterates its IndexMap, a std::unordered_map<std::string, ...> whose keys were inserted in
	// the order barcodes first had a pair stored: same container, same insertion order, same iteration order
	std::vector<uint32_t> present;
	for (const auto& kv : range_of)
		present.push_back(kv.first);
	std::sort(present.begin(), present.end(), [&](uint32_t a, uint32_t b) {
		const uint64_t fa = a < in.first_stored.size() ? in.first_stored[a] : UINT64_MAX;
		const uint64_t fb = b < in.first_stored.size() ? in.first_stored[b] : UINT64_MAX;
		return fa != fb ? fa < fb : a < b;
	});
	std::unordered_map<std::string, uint32_t> imap_shadow;
	for (uint32_t b : present)
		imap_shadow.emplace(in.bc_name[b], b);
	std::vector<uint32_t> bc_order;
	for (const auto& kv : imap_shadow)
		bc_order.push_back(kv.second);

	auto length_of = [&](uint32_t contig) { return (unsigned)in.to_length.at(in.ct_name[contig]); };
	auto in_mult_range = [&](uint32_t b) { return !(in.bc_mult[b] < params.min_mult || in.bc_mult[b] > params.max_mult); };

	// ---- calcDistSamples (DistanceEst.h:101-172)
	std::cout << "\n\t=> Measuring intra-contig distances / shared barcodes... " << stamp();
	std::unordered_map<std::string, DistSample> distSamples;
	for (uint32_t b : bc_order) {
		if (!in_mult_range(b))
			continue;
		const auto rg = range_of[b];
		for (size_t i = rg.first; i < rg.second; ++i) {
			const uint32_t row = rows[i], contig = in.im_ct[row];
			for (int isHead = 0; isHead < 2; ++isHead) {
				const int readPairs = (int)(isHead ? in.im_h[row] : in.im_t[row]);
				if (readPairs < params.min_reads)
					continue;
				const unsigned l = length_of(contig);
				if (l < two_e)
					continue;
				DistSample& s = distSamples[in.ct_name[contig]];
				s.distance = l - 2 * params.end_length;
				if (isHead)
					s.barcodesHead++;
				else
					s.barcodesTail++;
				const bool foundOther = (int)(isHead ? in.im_t[row] : in.im_h[row]) >= params.min_reads;
				if (foundOther && isHead) {
					s.barcodesIntersect++;
					s.barcodesUnion++;
				} else if (!foundOther) {
					s.barcodesUnion++;
				}
			}
		}
	}
	// ---- writeDistSamplesTSV (DistanceEst.h:501-535)
	std::cout << "\n\t=> Writing intra-contig distance samples to TSV... " << stamp();
	if (!params.dist_samples_tsv.empty()) {
		std::ofstream out(params.dist_samples_tsv.c_str());
		out << "contig_id" << '\t' << "distance" << '\t' << "barcodes_head" << '\t' << "barcodes_tail" << '\t' << "barcodes_union" << '\t'
		    << "barcodes_intersect" << '\n';
		for (const auto& it : distSamples)
			out << it.first << '\t' << it.second.distance << '\t' << it.second.barcodesHead << '\t' << it.second.barcodesTail << '\t'
			    << it.second.barcodesUnion << '\t' << it.second.barcodesIntersect << '\n';
	}
	// ---- buildJaccardToDist (DistanceEst.h:181-189): the first sample inserted under a Jaccard value stays
	std::cout << "\n\t=> Building Jaccard to distance map... " << stamp();
	std::map<double, DistSample> jaccardToDist;
	for (const auto& it : distSamples)
		jaccardToDist.insert(std::make_pair(double(it.second.barcodesIntersect) / it.second.barcodesUnion, it.second));

	// ---- buildPairToBarcodeStats (DistanceEst.h:220-334), for the contig pairs that are edges of g
	std::cout << "\n\t=> Calculating barcode stats for scaffold pairs... " << stamp();
	std::unordered_map<std::string, uint32_t> contig_of; // name -> first contig with that name
	for (uint32_t i = 0; i < in.ct_name.size(); ++i)
		contig_of.emplace(in.ct_name[i], i);
	struct PairStats
	{
		bool present = false;
		unsigned intersect[4] = { 0, 0, 0, 0 };
	};
	std::unordered_map<uint64_t, size_t> edge_of; // (contig a << 32 | contig b) -> edge index
	std::vector<PairStats> pstats(g.edges.size());
	for (size_t e = 0; e < g.edges.size(); ++e)
		edge_of[((uint64_t)contig_of.at(g.vid[g.edges[e].u]) << 32) | contig_of.at(g.vid[g.edges[e].v])] = e;
	std::vector<unsigned> end_barcodes(2 * in.ct_name.size(), 0); // contigEndToBarcodeCount[2 * contig + isHead]
	std::vector<std::pair<uint32_t, int>> valid; // (contig, isHead) entries of one barcode that pass validBarcodeMapping
	for (uint32_t b : bc_order) {
		if (!in_mult_range(b))
			continue;
		const auto rg = range_of[b];
		valid.clear();
		for (size_t i = rg.first; i < rg.second; ++i) {
			const uint32_t row = rows[i], contig = in.im_ct[row];
			if (length_of(contig) < two_e)
				continue;
			for (int isHead = 0; isHead < 2; ++isHead)
				if ((int)(isHead ? in.im_h[row] : in.im_t[row]) >= params.min_reads) {
					valid.emplace_back(contig, isHead);
					end_barcodes[2 * (size_t)contig + isHead]++;
				}
		}
		for (const auto& e1 : valid)
			for (const auto& e2 : valid) {
				if (in.lexrank[e1.first] >= in.lexrank[e2.first])
					continue; // id1 > id2 is skipped upstream; id1 == id2 is never an edge
				auto it = edge_of.find(((uint64_t)e1.first << 32) | e2.first);
				if (it == edge_of.end())
					continue;
				PairStats& ps = pstats[it->second];
				ps.present = true;
				ps.intersect[2 * (e1.second ? 0 : 1) + (e2.second ? 0 : 1)]++;
			}
	}
	auto stats_of = [&](size_t e) {
		const Edge& ed = g.edges[e];
		BarcodeStats st;
		const int i = ed.orientation;
		st.barcodesIntersect = pstats[e].intersect[i];
		const uint32_t a = contig_of.at(g.vid[ed.u]), b = contig_of.at(g.vid[ed.v]);
		const unsigned c1 = end_barcodes[2 * (size_t)a + ((i == 0 || i == 1) ? 1 : 0)];
		if (c1 == 0)
			return st;
		st.barcodes1 = c1;
		const unsigned c2 = end_barcodes[2 * (size_t)b + ((i == 0 || i == 2) ? 1 : 0)];
		if (c2 == 0)
			return st;
		st.barcodes2 = c2;
		st.barcodesUnion = st.barcodes1 + st.barcodes2 - st.barcodesIntersect;
		return st;
	};

	// ---- addEdgeDistances / estimateDistance (DistanceEst.h:337-430; closestKeys: Common/MapUtil.h:50-95)
	std::cout << "\n\t=> Adding edge distances... " << stamp();
	std::vector<std::pair<double, unsigned>> J; // (jaccard, distance), ascending
	for (const auto& it : jaccardToDist)
		J.emplace_back(it.first, it.second.distance);
	if (!J.empty())
		for (size_t e = 0; e < g.edges.size(); ++e) {
			if (!pstats[e].present)
				continue;
			const BarcodeStats st = stats_of(e);
			if (st.barcodesUnion == 0)
				continue;
			const double jac = double(st.barcodesIntersect) / st.barcodesUnion;
			// closestKey
			size_t it = std::lower_bound(J.begin(), J.end(), jac, [](const std::pair<double, unsigned>& x, double k) { return x.first < k; }) -
			            J.begin();
			size_t first;
			if (it == 0)
				first = 0;
			else if (it == J.size())
				first = J.size() - 1;
			else
				first = fabs(jac - J[it - 1].first) > fabs(jac - J[it].first) ? it : it - 1;
			size_t last = first + 1;
			for (size_t count = 1; count < params.dist_bin_size; ++count) {
				if (first == 0 && last == J.size())
					break;
				if (first == 0)
					++last;
				else if (last == J.size())
					--first;
				else if (fabs(jac - J[first - 1].first) < fabs(jac - J[last].first))
					--first;
				else
					++last;
			}
			std::vector<unsigned> distances;
			for (size_t i = first; i < last; ++i)
				distances.push_back(J[i].second);
			std::sort(distances.begin(), distances.end());
			Edge& ed = g.edges[e];
			ed.minDist = (int)floor(quantile(distances, 0.01));
			ed.dist = (int)round(quantile(distances, 0.5));
			ed.maxDist = (int)ceil(quantile(distances, 0.99));
			ed.jaccard = (float)jac;
		}
	// ---- writeDistTSV (DistanceEst.h:433-493)
	if (!params.dist_tsv.empty()) {
		std::cout << "\n\t=> Writing distance estimates to TSV... " << stamp();
		std::ofstream out(params.dist_tsv.c_str());
		out << "contig1" << '\t' << "contig2" << '\t' << "min_dist" << '\t' << "dist" << '\t' << "max_dist" << '\t' << "barcodes1" << '\t'
		    << "barcodes2" << '\t' << "barcodes_union" << '\t' << "barcodes_intersect" << '\n';
		for (size_t e = 0; e < g.edges.size(); ++e) {
			if (!pstats[e].present)
				continue;
			const Edge& ed = g.edges[e];
			const BarcodeStats st = stats_of(e);
			const bool sense1 = ed.orientation < 2, sense2 = ed.orientation % 2;
			const std::string &id1 = g.vid[ed.u], &id2 = g.vid[ed.v];
			for (int pass = 0; pass < 2; ++pass) {
				if (pass == 0)
					out << id1 << (sense1 ? '-' : '+') << '\t' << id2 << (sense2 ? '-' : '+') << '\t';
				else
					out << id2 << (sense2 ? '+' : '-') << '\t' << id1 << (sense1 ? '+' : '-') << '\t';
				if (ed.jaccard >= 0)
					out << ed.minDist << '\t' << ed.dist << '\t' << ed.maxDist << '\t';
				else
					out << "NA" << '\t' << "NA" << '\t' << "NA" << '\t';
				if (pass == 0)
					out << st.barcodes1 << '\t' << st.barcodes2;
				else
					out << st.barcodes2 << '\t' << st.barcodes1;
				out << '\t' << st.barcodesUnion << '\t' << st.barcodesIntersect << '\n';
			}
		}
	}
}

} // namespace

int main(int argc, char** argv)
{
	printf("Reading user inputs...\n");
	bool arcsOnly = false, arksOnly = false, dieflag = false;
	if (const char* s = getenv("ARKS_GPUS"))
		params.gpus = atoi(s);
	for (int c; (c = getopt_long(argc, argv, shortopts, longopts, NULL)) != -1;) {
		std::istringstream arg(optarg != NULL ? optarg : "");
		switch (c) {
		case 'u': arg >> params.multfile; break;
		case 'k': arg >> params.k_value; arksOnly = true; break;
		case 'j': arg >> params.j_index; arksOnly = true; break;
		case 't': arg >> params.threads; arksOnly = true; break;
		case '?': dieflag = true; break;
		case 'f': arg >> params.file; break;
		case 'a': arg >> params.fofName; break;
		case 'B': arg >> params.dist_bin_size; break;
		case 's': arg >> params.seq_id; arcsOnly = true; break;
		case 'c': arg >> params.min_reads; break;
		case 'P': params.output_pair = true; break;
		case 'D': params.dist_est = true; break;
		case 'l': arg >> params.min_links; break;
		case 'z': arg >> params.min_size; break;
		case 'b': arg >> params.base_name; break;
		case 'g': arg >> params.dist_graph_name; break;
		case OPT_TSV: arg >> params.tsv_name; break;
		case OPT_GAP: arg >> params.gap; break;
		case OPT_BARCODE_COUNTS: arg >> params.barcode_counts_name; break;
		case OPT_SAMPLES_TSV: arg >> params.dist_samples_tsv; break;
		case OPT_DIST_TSV: arg >> params.dist_tsv; break;
		case OPT_NO_DIST_EST: params.dist_est = false; break;
		case OPT_DIST_MEDIAN: params.dist_upper = false; break;
		case OPT_DIST_UPPER: params.dist_upper = true; break;
		case OPT_ARKS_METHOD: params.arks = true; break;
		case OPT_BX: break;
		case OPT_GPUS: arg >> params.gpus; break;
		case OPT_TWO_PASS: params.two_pass = true; break;
		case OPT_CUT:
			arg >> params.cut;
			arksOnly = true;
			break;
		case OPT_CUT_MIN:
			arg >> params.cut_min;
			arksOnly = true;
			break;
		case 'm': {
			std::string a, b;
			std::getline(arg, a, '-');
			std::getline(arg, b);
			std::stringstream ss;
			ss << a << "\t" << b;
			ss >> params.min_mult >> params.max_mult;
		} break;
		case 'd': arg >> params.max_degree; break;
		case 'e': arg >> params.end_length; break;
		case 'r': arg >> params.error_percent; break;
		case 'v': ++params.verbose; break;
		case OPT_HELP: std::cout << USAGE; exit(EXIT_SUCCESS);
		case OPT_VERSION: std::cout << PROGRAM " " PACKAGE_VERSION "\n"; exit(EXIT_SUCCESS);
		}
		if (optarg != NULL && c != 'm' && (!arg.eof() || arg.fail())) {
			std::cerr << PROGRAM ": invalid option: `-" << (char)c << optarg << "'\n";
			exit(EXIT_FAILURE);
		}
	}
	if ((params.arks && arcsOnly) || (!params.arks && arksOnly)) {
		std::cerr << PROGRAM ": error: You specified an option that does not match with method "
		                     "choosen.\nCheck --help for method specific options.\n";
		dieflag = true;
	}
	std::vector<std::string> filenames(argv + optind, argv + argc);
	if (params.fofName.empty() && filenames.empty()) {
		std::cerr << PROGRAM ": error: specify input (SAM/BAM file(s) or chromium reads) or a list of "
		                     "files with -a option\n";
		dieflag = true;
	}
	bool stdIn = !filenames.empty() && filenames[0] == "/dev/stdin";
	if (!params.file.empty())
		assert_readable(params.file);
	if (!params.fofName.empty())
		assert_readable(params.fofName);
	for (const auto& f : filenames)
		assert_readable(f);
	for (const auto& f : read_fof(params.fofName))
		filenames.push_back(f);
	bool alignmentFiles = false;
	if (!stdIn && !check_same_format(filenames, alignmentFiles)) {
		std::cerr << "Input files must be all alignment or all read files." << params.file << ". Exiting... \n";
		dieflag = true;
	}
	if (!stdIn && !(alignmentFiles ^ params.arks)) {
		std::cerr << "File type must be compatible with the method. (BAM/SAM for ARCS) or (Read "
		             "file for ARKS (--arks)). Exiting... \n";
		dieflag = true;
	}
	if (!params.arks)
		params.gpus = 1; // alignment mode: the tallies are made on the host; one GPU pairs them
	{
		std::ifstream g(params.file.c_str());
		if (!g.good() && params.arks) {
			std::cerr << "Cannot find [-f] scaffold file which is required for --arks" << params.file << ". Exiting... \n";
			dieflag = true;
		}
	}
	if (params.arks && (params.k_value < ARKS_MIN_K || params.k_value > ARKS_MAX_K)) {
		std::cerr << PROGRAM ": error: -k must be in [" << ARKS_MIN_K << ", " << ARKS_MAX_K << "] in this build.\n";
		dieflag = true;
	}
	if (params.base_name.empty()) { // Arcs.cpp:2139-2166
		std::ostringstream fn;
		if (params.arks)
			fn << params.file << ".scaff"
			   << "_arks"
			   << "_c" << params.min_reads << "_k" << params.k_value << "_j" << params.j_index << "_l" << params.min_links << "_d"
			   << params.max_degree << "_e" << params.end_length << "_r" << params.error_percent;
		else
			fn << params.file << ".scaff"
			   << "_arcs"
			   << "_s" << params.seq_id << "_c" << params.min_reads << "_l" << params.min_links << "_d" << params.max_degree << "_e"
			   << params.end_length << "_r" << params.error_percent;
		params.base_name = fn.str();
	}
	if (params.dist_graph_name.empty())
		params.dist_graph_name = params.base_name + ".dist.gv";
	if (params.tsv_name.empty())
		params.tsv_name = params.base_name + "_main.tsv";
	if (dieflag) {
		std::cerr << "Try " << PROGRAM << " --help for more information.\n";
		exit(EXIT_FAILURE);
	}
	if (params.gpus < 1)
		params.gpus = 1;
	if (params.gpus > (int)arks_host::kMaxShards) {
		std::cerr << PROGRAM ": error: --gpus must be at most " << arks_host::kMaxShards << ".\n";
		exit(EXIT_FAILURE);
	}
	printf("%s\n", "Finished reading user inputs...entering runArcs()...");

	// ---- runArcs banner (Arcs.cpp:1818-1843)
	std::cout << "Running: " << PROGRAM << " " << PACKAGE_VERSION << (params.arks ? "\nARKS" : "\nARCS") << " method\n pid " << ::getpid() << "\n -c "
	          << params.min_reads << "\n -d " << params.max_degree << "\n -e " << params.end_length << "\n -l " << params.min_links
	          << "\n -m " << params.min_mult << '-' << params.max_mult << "\n -r " << params.error_percent << "\n -v "
	          << params.verbose << "\n -z " << params.min_size << "\n --gap=" << params.gap << "\n -k " << params.k_value << "\n -j "
	          << params.j_index << "\n -t " << params.threads << "\n --gpus " << params.gpus << "\n -b "
	          << maybeNA(params.base_name) << "\n -g " << maybeNA(params.dist_graph_name)
	          << "\n --barcode-counts=" << maybeNA(params.barcode_counts_name) << "\n --tsv=" << maybeNA(params.tsv_name) << "\n -a "
	          << maybeNA(params.fofName) << "\n -f " << maybeNA(params.file) << "\n -u " << maybeNA(params.multfile) << '\n';
	for (const auto& f : filenames)
		std::cout << ' ' << f << '\n';
	std::cout.flush();
	const double t_start = now();

	Barcodes bc;
	Contigs ct;
	std::vector<Gpu> gpus(params.gpus);
	std::thread comm_init; // sets up the NCCL communicator of the pair-link merge while the reads are mapped
	int comm_rc = ARKS_OK;
	double t_index0 = now(), t_index1 = t_index0, t_map0 = t_index0, t_map1 = t_index0, t_gpu_init = 0, t_draft0 = 0, t_draft1 = 0, t_ctx0 = t_index0, t_ctx1 = t_index0;
	bool draft_fast = false;
	if (!params.arks) {
		// ---- ARCS alignment mode (runArcs :1859-1871): contig sizes from -f (getScaffSizes :549-568) and/or
		// the SAM headers, tallies from the alignments, then the same pair-link / graph path as ARKS
		std::unordered_map<std::string, uint32_t> contig_id;
		if (!params.file.empty()) {
			std::cout << "\n=> Getting scaffold sizes... " << stamp();
			SeqReader rd(params.file);
			if (!rd.ok())
				die("error: cannot open " + params.file);
			SeqRecord r;
			int counter = 0;
			while (rd.read(r) >= 0) {
				r.truncate_at_nul();
				counter++;
				intern_contig(ct, contig_id, r.name, (int)r.seq.size(), true);
			}
			if (params.verbose)
				std::cout << "Saw " << counter << " sequences.\n";
		}
		std::cout << "\n=> Reading alignment files... " << stamp();
		t_map0 = now();
		AlignRows rows;
		for (const auto& f : filenames) {
			if (params.verbose)
				std::cout << "Reading alignments: " << f << std::endl;
			read_alignments(f, bc, ct, contig_id, rows);
		}
		g_first_stored = rows.first_add;
		Gpu& g = gpus[0];
		if (arks_create(0, 32, 1024, &g.h) != ARKS_OK)
			die(std::string("error: cannot initialise GPU 0: ") + arks_last_error(nullptr));
		std::vector<uint32_t> rb, rc, rh, rt;
		for (const auto& kv : rows.ht) {
			rb.push_back((uint32_t)(kv.first >> 32));
			rc.push_back((uint32_t)kv.first);
			rh.push_back(kv.second.first);
			rt.push_back(kv.second.second);
			// pairContigs looks the multiplicity up with operator[] (Arcs.cpp:1389): a barcode that was only
			// ever seen on secondary / supplementary alignments becomes a key with multiplicity 0
			bc.counted[rb.back()] = 1;
		}
		if (!rb.empty())
			ck(g.h, arks_imap_add(g.h, rb.data(), rc.data(), rh.data(), rt.data(), rb.size()), "arks_imap_add");
		t_map1 = now();
	} else {
	// the CUDA context is created on a helper thread while the multiplicity file and the draft are parsed
	// (only the GPUs this run uses are made visible: the CUDA runtime initialises every visible device of the node,
	// which costs seconds on an 8-GPU box)
	if (!getenv("CUDA_VISIBLE_DEVICES")) {
		std::string vis;
		for (int d = 0; d < (getenv("ARKS_GPUS_SAME_DEVICE") ? 1 : params.gpus); ++d)
			vis += (d ? "," : "") + std::to_string(d);
		setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 0);
	}
	t_ctx0 = now();
	t_ctx1 = t_ctx0;
	std::thread gpu_init([&t_ctx1] {
		std::vector<std::thread> th;
		for (int d = 1; d < params.gpus && !getenv("ARKS_GPUS_SAME_DEVICE"); ++d)
			th.emplace_back([d] { arks_device_init(d); });
		arks_device_init(0);
		for (auto& t : th)
			t.join();
		t_ctx1 = now();
	});
	// ---- barcode multiplicities
	const bool have_multfile = !params.multfile.empty();
	// `goodmult = mult > min || mult < max` (Arcs.cpp:1267) can only be false for an inverted range
	const bool need_first_pass = !have_multfile && (params.two_pass || params.min_mult >= params.max_mult);
	std::cout << "\n=>Preprocessing: Gathering barcode multiplicity information..." << stamp();
	if (have_multfile) {
		load_multfile(params.multfile, bc);
	} else if (need_first_pass) {
		// readBarcodes (Arcs.cpp:481-547)
		std::string scratch;
		for (const auto& f : filenames) {
			SeqReader rd(open_reads(f), std::string());
			SeqRecord r;
			while (rd.read(r) > 0) {
				r.truncate_at_nul();
				count_barcode(r, bc, scratch);
			}
		}
		if (params.verbose)
			std::cout << "Saw " << bc.name.size() << " distinct barcode." << std::endl;
	} else {
		std::cout << "Multiplicity information is being formed from reads while they are mapped (single pass)." << std::endl;
	}
	const bool mult_known = have_multfile || need_first_pass;

	// ---- contigs: getContigKmers (Arcs.cpp:1021-1129) on the GPU
	std::cout << "\n=>Preprocessing: Gathering draft information..." << stamp() << "\n";
	std::vector<char> end_bases_vec;
	std::unique_ptr<char[]> end_bases_raw; // the fast path fills an uninitialised buffer from several threads
	const char* end_bases_ptr = nullptr;
	uint64_t end_bases_n = 0;
	std::vector<uint64_t> end_off(1, 0);
	std::vector<uint32_t> end_conreci;
	t_draft0 = now();
	{
		std::unordered_map<std::string, uint32_t> first;
		// one kept contig: bookkeeping + the extent of its two ends (getContigKmers, Arcs.cpp:1056-1091)
		auto keep = [&](const std::string& name, int len) -> int {
			const uint32_t i = (uint32_t)ct.name.size();
			ct.name.push_back(name);
			ct.length.push_back(len);
			ct.to_length[name] = len;
			auto it = first.find(name);
			ct.first_of_name.push_back(it == first.end() ? i : it->second);
			if (it == first.end())
				first.emplace(name, i);
			int cut = params.end_length;
			if (cut == 0 || len <= cut * 2)
				cut = len / 2;
			end_off.push_back(end_off.back() + (uint64_t)cut);
			end_conreci.push_back(2 * i + 1);
			end_off.push_back(end_off.back() + (uint64_t)cut);
			end_conreci.push_back(2 * i + 2);
			return cut;
		};
		const int threads = (int)std::max(1u, std::thread::hardware_concurrency());
		arks_host::MappedFasta mf;
		if (!getenv("ARKS_NO_FAST_FASTA") && mf.open(params.file, threads)) {
			// plain strict FASTA: every core lists records, then copies the ends (host/fasta_fast.h)
			draft_fast = true;
			const auto& recs = mf.records();
			std::vector<uint32_t> kept;   // record index of every kept contig
			std::vector<uint32_t> cuts;
			for (size_t r = 0; r < recs.size(); ++r) {
				ct.total++;
				if (recs[r].seq_len > (size_t)INT_MAX)
					die("error: contig longer than 2^31 bases");
				const int len = (int)recs[r].seq_len;
				if (len < params.min_size) {
					ct.skipped++;
					continue;
				}
				kept.push_back((uint32_t)r);
				cuts.push_back((uint32_t)keep(std::string(recs[r].name, recs[r].name_n), len));
			}
			end_bases_n = end_off.back();
			end_bases_raw.reset(new char[end_bases_n + 64]);
			end_bases_ptr = end_bases_raw.get();
			const size_t nt = std::max<size_t>(1, std::min<size_t>((size_t)threads, kept.size() / 64 + 1));
			auto work = [&](size_t t) {
				for (size_t c = kept.size() * t / nt; c < kept.size() * (t + 1) / nt; ++c) {
					const arks_host::FastaRecord& r = recs[kept[c]];
					arks_host::MappedFasta::copy_bases(r, 0, cuts[c], end_bases_raw.get() + end_off[2 * c]);
					arks_host::MappedFasta::copy_bases(r, r.seq_len - cuts[c], cuts[c], end_bases_raw.get() + end_off[2 * c + 1]);
				}
			};
			std::vector<std::thread> th;
			for (size_t t = 1; t < nt; ++t)
				th.emplace_back(work, t);
			work(0);
			for (auto& x : th)
				x.join();
		} else {
			ct = Contigs();
			end_off.assign(1, 0);
			end_conreci.clear();
			SeqReader rd(params.file);
			if (!rd.ok())
				die("error: cannot open " + params.file);
			SeqRecord r;
			while (rd.read(r) >= 0) {
				ct.total++;
				r.truncate_at_nul();
				const int len = (int)r.seq.size();
				if (len < params.min_size) {
					ct.skipped++;
					continue;
				}
				const int cut = keep(r.name, len);
				end_bases_vec.insert(end_bases_vec.end(), r.seq.begin(), r.seq.begin() + cut);
				end_bases_vec.insert(end_bases_vec.end(), r.seq.end() - cut, r.seq.end());
			}
			end_bases_ptr = end_bases_vec.data();
			end_bases_n = end_bases_vec.size();
		}
	}
	t_draft1 = now();
	if (params.verbose)
		std::cerr << "Number of contigs:" << ct.name.size() << "\nSize of Contig Array:" << ct.name.size() * 2 + 1 << std::endl;

	std::cout << "\n=>Storing Kmers from Contig ends... " << stamp() << std::endl;
	gpu_init.join();
	arks_index_stats ist{};
	t_index0 = now();
	{
		// hits on a contig whose name occurred earlier are tallied under the first one (imap is keyed by name)
		std::vector<uint32_t> remap(2 * ct.name.size() + 1, 0);
		bool any = false;
		for (uint32_t i = 0; i < ct.name.size(); ++i) {
			remap[2 * i + 1] = 2 * ct.first_of_name[i] + 1;
			remap[2 * i + 2] = 2 * ct.first_of_name[i] + 2;
			any |= ct.first_of_name[i] != i;
		}
		// every GPU builds the same index from the same host buffers, all at once: one host thread per GPU
		std::vector<arks_index_stats> ists(params.gpus);
		std::vector<double> t_create(params.gpus, 0.0);
		auto build = [&](int d) {
			Gpu& g = gpus[d];
			bind_thread_to_gpu(d);
			// ARKS_GPUS_SAME_DEVICE=1 (tests): every shard on device 0
			const int dev = getenv("ARKS_GPUS_SAME_DEVICE") ? 0 : d;
			const double tc0 = now();
			int rc = arks_create(dev, params.k_value, end_bases_n + 64, &g.h);
			t_create[d] = now() - tc0;
			if (rc != ARKS_OK)
				die(std::string("error: cannot initialise GPU ") + std::to_string(d) + ": " + arks_last_error(nullptr));
			if (!end_conreci.empty())
				ck(g.h, arks_index_add(g.h, end_bases_ptr, end_off.data(), end_conreci.data(), (uint32_t)end_conreci.size()),
				    "arks_index_add");
			ck(g.h, arks_index_finalize(g.h, &ists[d]), "arks_index_finalize");
			if (any)
				ck(g.h, arks_set_conreci_remap(g.h, remap.data(), (uint32_t)remap.size()), "arks_set_conreci_remap");
		};
		std::vector<std::thread> th;
		for (int d = 1; d < params.gpus; ++d)
			th.emplace_back(build, d);
		build(0);
		for (auto& t : th)
			t.join();
		ist = ists[0];
		t_gpu_init = *std::max_element(t_create.begin(), t_create.end());
	}
	t_index1 = now();
	// the NCCL communicator of the merge takes seconds to set up: done on a helper thread while the reads are mapped
	if (params.gpus > 1)
		comm_init = std::thread([&] {
			std::vector<arks_handle*> hs;
			for (auto& g : gpus)
				hs.push_back(g.h);
			comm_rc = arks_comm_init_local(hs.data(), params.gpus);
		});
	std::vector<char>().swap(end_bases_vec);
	end_bases_raw.reset();
	if (params.verbose)
		printf("%s %zu\n%s %zu\n%s %zu\n%s %llu\n%s %llu\n%s %llu\n%s %llu\n%s %llu\n%s %llu\n",
		    "Total number of contigs in draft genome: ", ct.total, "Total valid contigs: ", ct.name.size(),
		    "Total skipped contigs: ", ct.skipped, "Total number of Kmers: ", (unsigned long long)ist.kmers_valid,
		    "Number Null Kmers: ", (unsigned long long)ist.kmers_null, "Number Kmers Recorded: ", (unsigned long long)ist.recorded,
		    "Number Kmer Collisions: ", (unsigned long long)ist.collisions,
		    "Number Times Kmers Removed (since duplicate in different contig): ", (unsigned long long)ist.removed,
		    "Number of unique kmers (only one contig): ", (unsigned long long)ist.unique);

	// ---- reads: chromiumRead (Arcs.cpp:1132-1351)
	std::cout << "\n=>Reading Chromium FASTQ file(s)... " << stamp() << std::endl;
	t_map0 = now();
	size_t skipped_unpaired = 0, emptybarcode = 0, invalidbarcode = 0, skipped_badmult = 0;
	size_t fast_blocks = 0;
	{
		arks_host::IngestConfig cfg;
		cfg.mult_known = mult_known;
		cfg.verbose = params.verbose != 0;
		cfg.min_mult = params.min_mult;
		cfg.max_mult = params.max_mult;
		arks_host::IngestCounters ictr;
		arks_host::PairSink sink;
		cfg.shards = (uint32_t)params.gpus;
		sink.add = [&](const std::string& s1, const std::string& s2, uint32_t id) {
			const std::string& name = bc.name[id];
			add_pair(gpus[arks_host::barcode_shard(name.data(), name.size(), cfg.shards)], s1.data(), s1.size(), s2.data(), s2.size(), id);
		};
		sink.submit = [&](const arks_host::PairBatch& b) {
			// the block is already a batch in pinned memory, laid out by shard: every GPU gets its part with one
			// asynchronous call, then the copies of all of them are awaited together
			for (uint32_t g = 0; g < b.n_shards; ++g)
				if (b.shard_pairs(g))
					ck(gpus[g].h, arks_map_pairs_begin(gpus[g].h, b.bases, b.shard_off(g), b.shard_bc(g), b.shard_pairs(g), params.j_index, nullptr),
					    "arks_map_pairs");
			for (uint32_t g = 0; g < b.n_shards; ++g)
				if (b.shard_pairs(g))
					ck(gpus[g].h, arks_map_pairs_end(gpus[g].h), "arks_map_pairs");
		};
		// block-parallel parsing (ingest.h) unless -D needs the pairs' input order or ARKS_PARSE_THREADS=0
		arks_host::ParallelIngestOptions popt;
		{
			// parser threads: the cores of the host minus the reader, the committer, the driver's threads; 4 at least
			// (the blocks they work on are in flight anyway), 24 at most (memory bandwidth, not cores, bounds it by then)
			const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
			popt.workers = (int)std::min(24u, std::max(4u, hw > 4 ? hw - 4 : 1u));
		}
		if (const char* e = getenv("ARKS_PARSE_THREADS"))
			popt.workers = atoi(e);
		// several GPUs: every block is dealt to all of them, so the blocks grow with their number (the per-GPU part
		// of a block stays about 4 MB and the committing thread issues as many calls per byte as with one GPU)
		popt.block_bytes *= (size_t)std::min(params.gpus, 8);
		if (const char* e = getenv("ARKS_PARSE_BLOCK_MB"))
			popt.block_bytes = (size_t)std::max(1, atoi(e)) << 20;
		const bool parallel = popt.workers > 0 && !params.dist_est;
		std::vector<void*> pinned;
		if (parallel) {
			const uint32_t cap_pairs = (uint32_t)(popt.block_bytes / 48 + 16);
			for (int i = 0; i < popt.workers + 2; ++i) {
				arks_host::PairBatch pb;
				void* p;
				if (arks_host_alloc(&p, popt.block_bytes + 4096) != ARKS_OK)
					die("error: cannot allocate pinned host memory");
				pb.bases = (char*)p;
				pinned.push_back(p);
				if (arks_host_alloc(&p, (2ull * cap_pairs + 1 + arks_host::kMaxShards) * 4) != ARKS_OK)
					die("error: cannot allocate pinned host memory");
				pb.off = (uint32_t*)p;
				pinned.push_back(p);
				if (arks_host_alloc(&p, cap_pairs * 4ull) != ARKS_OK)
					die("error: cannot allocate pinned host memory");
				pb.bc = (uint32_t*)p;
				pinned.push_back(p);
				pb.cap_bases = popt.block_bytes;
				pb.cap_pairs = cap_pairs;
				popt.slots.push_back(pb);
			}
		}
		for (const auto& f : filenames) {
			if (params.verbose)
				std::cout << "Reading chrom " << f << std::endl;
			bool counting = !mult_known; // readBarcodes stops at the first record with l <= 0
			if (parallel) {
				{
					std::ifstream probe(f.c_str());
					if (!probe.good()) {
						std::cerr << "File " << f << " cannot be opened." << std::endl;
						exit(1);
					}
				}
				std::cerr << "File " << f << " opened." << std::endl;
				size_t nb = 0;
				if (!arks_host::ingest_parallel_blocks(f, bc, cfg, counting, ictr, sink, popt, &nb, params.cut ? open_reads(f) : nullptr)) {
					std::cerr << "File " << f << " cannot be opened." << std::endl;
					exit(1);
				}
				fast_blocks += nb;
				continue;
			}
			SeqReader rd(open_reads(f), std::string());
			std::cerr << "File " << f << " opened." << std::endl;
			arks_host::ingest_sequential(rd, bc, cfg, counting, ictr, sink);
		}
		for (auto& g : gpus)
			flush_batch(g);
		for (void* p : pinned)
			arks_host_free(p);
		skipped_unpaired = ictr.skipped_unpaired;
		emptybarcode = ictr.emptybarcode;
		invalidbarcode = ictr.invalidbarcode;
		skipped_badmult = ictr.skipped_badmult;
	}
	arks_map_stats mst{};
	for (auto& g : gpus) {
		arks_map_stats s{};
		ck(g.h, arks_map_get_stats(g.h, &s), "arks_map_get_stats");
		uint64_t* a = reinterpret_cast<uint64_t*>(&mst);
		const uint64_t* b = reinterpret_cast<const uint64_t*>(&s);
		for (size_t i = 0; i < sizeof(mst) / 8; ++i)
			a[i] += b[i];
	}
	t_map1 = now();
	if (params.verbose) {
		printf("Stored read pairs: %llu\nSkipped invalid read pairs: %llu\nSkipped unpaired reads: %zu\nSkipped reads pairs without a "
		       "good contig: %llu\n",
		    (unsigned long long)mst.pairs_stored, (unsigned long long)(mst.pairs_invalid + skipped_badmult), skipped_unpaired,
		    (unsigned long long)(mst.pairs_nogood + skipped_badmult));
		printf("Total valid kmers: %llu\nNumber invalid kmers: %llu\nNumber of kmers found in ContigKmap: %llu\nNumber of kmers "
		       "recorded in Ktrack: %llu\nNumber of kmers found in ContigKmap but duplicate: %llu\nNumber of reads passing jaccard "
		       "threshold: %llu\nNumber of reads failing jaccard threshold: %llu\n",
		    (unsigned long long)mst.kmers_valid, (unsigned long long)mst.kmers_invalid, (unsigned long long)mst.found,
		    (unsigned long long)mst.recorded, (unsigned long long)mst.dups, (unsigned long long)mst.reads_pass,
		    (unsigned long long)mst.reads_fail);
		if (emptybarcode > 0)
			printf("WARNING:: Your chromium read file has %zu readpairs that have an empty barcode.", emptybarcode);
		if (invalidbarcode > 0)
			printf("WARNING:: Your chromium read file has %zu read pairs that have barcodes not in the barcode multiplicity file.",
			    invalidbarcode);
		const double kmers = (double)(mst.kmers_valid + mst.kmers_invalid);
		printf("GPU mapping: %.3f s wall (parse + H2D + kernels), %.3e read k-mers/s end to end; index build %.3f s; %zu blocks parsed in "
		       "parallel\n",
		    t_map1 - t_map0, kmers / std::max(1e-9, t_map1 - t_map0), t_index1 - t_index0, fast_blocks);
	}

	} // ARKS

	// ---- pairContigs (Arcs.cpp:1378-1435) on the GPU
	std::cout << "\n=> Pairing scaffolds... " << stamp();
	const uint32_t n_bc = (uint32_t)bc.name.size();
	std::vector<int32_t> mult(n_bc);
	for (uint32_t i = 0; i < n_bc; ++i)
		mult[i] = bc.counted[i] ? bc.mult[i] : INT32_MIN; // a barcode that is not a key of indexMultMap never had its pairs stored
	// rank of contig names under std::string '<' (PairMap's key order)
	const uint32_t n_ct = (uint32_t)ct.name.size();
	std::vector<uint32_t> order(n_ct), lexrank(n_ct);
	for (uint32_t i = 0; i < n_ct; ++i)
		order[i] = i;
	std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return ct.name[a] < ct.name[b] || (ct.name[a] == ct.name[b] && a < b); });
	for (uint32_t r = 0, rank = 0; r < n_ct; ++r) {
		if (r > 0 && ct.name[order[r]] != ct.name[order[r - 1]])
			rank = r;
		lexrank[order[r]] = rank;
	}
	struct PairRow
	{
		uint32_t a, b, c[4];
	};
	std::vector<PairRow> pmap;
	// imap rows (for the TSV): barcode, contig, head, tail
	std::vector<uint32_t> im_bc, im_ct, im_h, im_t;
	const double t_links0 = now();
	{
		// every GPU pairs the contigs of ITS barcodes (one host thread per GPU) ...
		std::vector<uint64_t> im_n(params.gpus, 0);
		std::vector<std::vector<uint32_t>> im(4 * (size_t)params.gpus);
		auto links = [&](int d) {
			Gpu& g = gpus[d];
			bind_thread_to_gpu(d);
			uint64_t n_rows = 0;
			ck(g.h, arks_imap_size(g.h, &n_rows), "arks_imap_size");
			for (int q = 0; q < 4; ++q)
				im[4 * d + q].resize(n_rows);
			if (n_rows)
				ck(g.h, arks_imap_export(g.h, im[4 * d].data(), im[4 * d + 1].data(), im[4 * d + 2].data(), im[4 * d + 3].data(), n_rows, &n_rows),
				    "arks_imap_export");
			im_n[d] = n_rows;
			ck(g.h, arks_pair_links(g.h, mult.data(), n_bc, params.min_mult, params.max_mult, params.min_reads, params.error_percent,
			            lexrank.data(), n_ct),
			    "arks_pair_links");
		};
		std::vector<std::thread> th;
		for (int d = 1; d < params.gpus; ++d)
			th.emplace_back(links, d);
		links(0);
		for (auto& t : th)
			t.join();
		for (int d = 0; d < params.gpus; ++d) {
			im_bc.insert(im_bc.end(), im[4 * d].begin(), im[4 * d].end());
			im_ct.insert(im_ct.end(), im[4 * d + 1].begin(), im[4 * d + 1].end());
			im_h.insert(im_h.end(), im[4 * d + 2].begin(), im[4 * d + 2].end());
			im_t.insert(im_t.end(), im[4 * d + 3].begin(), im[4 * d + 3].end());
		}
	}
	const double t_links1 = now();
	if (params.gpus > 1) {
		// ... and since barcodes are disjoint across GPUs the link maps add up key by key: one exchange over
		// NCCL (all-gather of the sorted keys, one all-reduce of the dense counters), after which GPU 0 holds the sum
		std::vector<arks_handle*> hs;
		for (auto& g : gpus)
			hs.push_back(g.h);
		if (comm_init.joinable())
			comm_init.join();
		ck(hs[0], comm_rc, "arks_comm_init_local");
		ck(hs[0], arks_merge_pmap(hs.data(), params.gpus), "arks_merge_pmap");
	}
	const double t_merge1 = now();
	{
		Gpu& g = gpus[0];
		uint64_t n = 0;
		ck(g.h, arks_pmap_size(g.h, &n), "arks_pmap_size");
		if (n) {
			void* p = nullptr;
			if (arks_host_alloc(&p, n * 24) != ARKS_OK)
				die("error: cannot allocate pinned host memory");
			uint32_t *a = (uint32_t*)p, *b = a + n, *c = b + n;
			ck(g.h, arks_pmap_export(g.h, a, b, c, n, &n), "arks_pmap_export");
			pmap.resize(n);
			for (uint64_t i = 0; i < n; ++i)
				pmap[i] = PairRow{ a[i], b[i], { c[4 * i], c[4 * i + 1], c[4 * i + 2], c[4 * i + 3] } };
			arks_host_free(p);
		}
	}
	const double t_export1 = now();
	// drop imap rows of barcodes that are not keys of indexMultMap (their pairs were never stored)
	{
		size_t w = 0;
		for (size_t i = 0; i < im_bc.size(); ++i)
			if (bc.counted[im_bc[i]]) {
				im_bc[w] = im_bc[i];
				im_ct[w] = im_ct[i];
				im_h[w] = im_h[i];
				im_t[w] = im_t[i];
				w++;
			}
		im_bc.resize(w);
		im_ct.resize(w);
		im_h.resize(w);
		im_t.resize(w);
	}

	// ARKS_DUMP_IMAP=FILE: the IndexMap as text, one "barcode contig H|T count" line per entry (both ends of
	// every contig a barcode touches, as after the zero-fill of Arcs.cpp:1309-1319), sorted -- for parity tests
	if (const char* dump = getenv("ARKS_DUMP_IMAP")) {
		std::vector<std::string> lines;
		for (size_t i = 0; i < im_bc.size(); ++i) {
			lines.push_back(bc.name[im_bc[i]] + "\t" + ct.name[im_ct[i]] + "\tH\t" + std::to_string(im_h[i]));
			lines.push_back(bc.name[im_bc[i]] + "\t" + ct.name[im_ct[i]] + "\tT\t" + std::to_string(im_t[i]));
		}
		std::sort(lines.begin(), lines.end());
		std::ofstream out(dump);
		for (const auto& l : lines)
			out << l << "\n";
	}

	if (params.output_pair) {
		std::cout << "\n=> Outputting Pairing information... " << stamp();
		std::ofstream out((params.base_name + "_pair.tsv").c_str());
		for (const auto& p : pmap)
			out << ct.name[p.a] << "\t" << ct.name[p.b] << "\t" << p.c[0] << "\t" << p.c[1] << "\t" << p.c[2] << "\t" << p.c[3] << std::endl;
	}

	// ---- createGraph (Arcs.cpp:1475-1526)
	std::cout << "\n=> Creating the graph... " << stamp();
	Graph g;
	{
		std::unordered_map<uint32_t, int> vmap; // by first-of-name contig index
		for (const auto& p : pmap) {
			unsigned max = 0, index = 0;
			for (unsigned i = 0; i < 4; ++i)
				if (p.c[i] > max) {
					max = p.c[i];
					index = i;
				}
			unsigned second = 0;
			for (unsigned i = 0; i < 4; ++i)
				if (p.c[i] != max && p.c[i] > second)
					second = p.c[i];
			if (check_significance((int)max, (int)(max + second))) {
				for (uint32_t cidx : { p.a, p.b })
					if (!vmap.count(cidx)) {
						vmap[cidx] = (int)g.vid.size();
						g.vid.push_back(ct.name[cidx]);
					}
				g.edges.push_back(Edge{ vmap[p.a], vmap[p.b], (int)index, (int)max });
			}
		}
	}

	// ---- calcDistanceEstimates (Arcs.cpp:1769-1807, Arcs/DistanceEst.h)
	if (params.dist_est) {
		std::cout << "\n=> Calculating distance estimates... " << stamp();
		DistInput in{ im_bc, im_ct, im_h, im_t, bc.name, bc.mult, ct.name, ct.to_length, lexrank, g_first_stored };
		calc_distance_estimates(in, g);
	}

	// ---- writePostRemovalGraph / removeDegreeNodes / writeGraph (Arcs.cpp:1549-1610)
	std::cout << "\n=> Writing graph file... " << stamp() << "\n";
	const std::string graphFile = params.base_name + "_original.gv";
	if (params.max_degree != 0) {
		std::cout << "      Deleting nodes with degree > " << params.max_degree << "... \n";
		std::vector<int> deg(g.vid.size(), 0);
		for (const auto& e : g.edges) {
			deg[e.u]++;
			deg[e.v]++;
		}
		std::vector<int> remap(g.vid.size(), -1);
		std::vector<std::string> nv;
		for (size_t i = 0; i < g.vid.size(); ++i)
			if (deg[i] <= params.max_degree) {
				remap[i] = (int)nv.size();
				nv.push_back(g.vid[i]);
			}
		std::vector<Edge> ne;
		for (const auto& e : g.edges)
			if (remap[e.u] >= 0 && remap[e.v] >= 0) {
				Edge n = e;
				n.u = remap[e.u];
				n.v = remap[e.v];
				ne.push_back(n);
			}
		g.vid.swap(nv);
		g.edges.swap(ne);
	} else {
		std::cout << "      Max Degree (-d) set to: " << params.max_degree << ". Will not delete any vertices from graph.\n";
	}
	std::cout << "      Writing graph file to " << graphFile << "...\n";
	{
		std::ofstream out(graphFile.c_str());
		if (!out)
			die("error: cannot write " + graphFile);
		out << "graph G {\n";
		for (size_t i = 0; i < g.vid.size(); ++i)
			out << i << " [id=" << g.vid[i] << "];\n";
		for (const auto& e : g.edges) {
			out << e.u << "--" << e.v << " [label=" << e.orientation << ", weight=" << e.weight;
			if (e.minDist != INT_MIN) // EdgePropertyWriter, Arcs/Arcs.h:204-211
				out << ", d=" << e.dist << ", maxd=" << e.maxDist;
			out << "];\n";
		}
		out << "}\n";
	}
	const double t_gv = now();

	// ---- createAbyssGraph / writeAbyssGraph (Arcs.cpp:1615-1672; Graph/DotIO.h:82-114)
	std::cout << "\n=> Creating the ABySS graph... " << stamp();
	std::cout << "\n=> Writing the ABySS graph file... " << stamp() << "\n";
	{
		// two vertices per contig ("name+" then "name-") in ContigToLength iteration order; an edge
		// u -> v also adds its complement v^ -> u^
		std::vector<std::string> vname;
		std::vector<int> vlen;
		std::unordered_map<std::string, size_t> vindex; // contig name -> index of its '+' vertex
		for (const auto& it : ct.to_length) {
			vindex[it.first] = vname.size();
			vname.push_back(it.first + "+");
			vlen.push_back(it.second);
			vname.push_back(it.first + "-");
			vlen.push_back(it.second);
		}
		struct Out
		{
			size_t to;
			int n;
			int d;
		};
		std::vector<std::vector<Out>> adj(vname.size());
		auto has_edge = [&](size_t u, size_t v) {
			for (const auto& o : adj[u])
				if (o.to == v)
					return true;
			return false;
		};
		for (const auto& e : g.edges) {
			const size_t u = vindex[g.vid[e.u]] + (e.orientation < 2 ? 1 : 0);
			const size_t v = vindex[g.vid[e.v]] + (e.orientation % 2 ? 1 : 0);
			if (has_edge(u, v)) {
				std::cerr << "error: Duplicate edge: \"" << vname[u] << "\" -> \"" << vname[v] << '"' << std::endl;
				exit(EXIT_FAILURE);
			}
			// ep.distance: the fixed gap, or under -D the estimate (INT_MAX where there is none), Arcs.cpp:1635-1648
			const int d = params.dist_est ? (params.dist_upper ? e.maxDist : e.dist) : (int)params.gap;
			adj[u].push_back(Out{ v, e.weight, d });
			const size_t uc = u ^ 1, vc = v ^ 1;
			if (!(vc == u && uc == v))
				adj[vc].push_back(Out{ uc, e.weight, d });
		}
		std::ofstream out(params.dist_graph_name.c_str());
		if (!out)
			die("error: cannot write " + params.dist_graph_name);
		out << "digraph arcs {\n";
		for (size_t i = 0; i < vname.size(); ++i)
			out << '"' << vname[i] << "\" [l=" << vlen[i] << "]\n";
		for (size_t u = 0; u < vname.size(); ++u)
			for (const auto& o : adj[u])
				out << '"' << vname[u] << "\" -> \"" << vname[o.to] << "\" [d=" << o.d << " e=" << std::fixed
				    << std::setprecision(1) << (float)params.gap << " n=" << o.n << "]\n";
		out << "}\n";
	}

	// ---- countBarcodes + writeTSV (Arcs.cpp:815-830,1709-1757)
	if (!params.tsv_name.empty()) {
		size_t barcodeCount = 0, n_keys = 0;
		for (uint32_t i = 0; i < n_bc; ++i)
			if (bc.counted[i]) {
				n_keys++;
				if (bc.mult[i] >= params.min_mult && bc.mult[i] <= params.max_mult)
					++barcodeCount;
			}
		std::vector<uint8_t> seen(n_bc, 0);
		size_t scaffold_end_barcodes = 0;
		for (uint32_t b : im_bc)
			if (!seen[b]) {
				seen[b] = 1;
				scaffold_end_barcodes++;
			}
		std::cout << "{ \"All_barcodes_unfiltered\":" << n_keys << ", \"All_barcodes_filtered\":" << barcodeCount
		          << ", \"Scaffold_end_barcodes\":" << scaffold_end_barcodes << ", \"Min_barcode_reads_threshold\":" << params.min_mult
		          << ", \"Max_barcode_reads_threshold\":" << params.max_mult << " }\n";
		std::cout << "\n=> Writing TSV file... " << stamp();
		// barcodes per scaffold end with count >= min_reads (a missing end is a zero-count entry, Arcs.cpp:1309-1319)
		std::vector<unsigned> per_end(2 * (size_t)n_ct, 0); // [2*contig + (head ? 0 : 1)]
		for (size_t i = 0; i < im_bc.size(); ++i) {
			if ((int)im_h[i] >= params.min_reads)
				per_end[2 * (size_t)im_ct[i]]++;
			if ((int)im_t[i] >= params.min_reads)
				per_end[2 * (size_t)im_ct[i] + 1]++;
		}
		std::ofstream f(params.tsv_name.c_str());
		if (!f)
			die("error: cannot write " + params.tsv_name);
		f << "U\tV\tBest_orientation\tShared_barcodes\tU_barcodes\tV_barcodes\tAll_barcodes\n";
		for (const auto& p : pmap) {
			const std::string& u = ct.name[p.a];
			const std::string& v = ct.name[p.b];
			unsigned max_counts = *std::max_element(p.c, p.c + 4);
			for (unsigned i = 0; i < 4; ++i) {
				if (p.c[i] == 0)
					continue;
				const bool usense = i < 2, vsense = i % 2;
				// barcodes_per_scaffold_end[(u, usense)] / [(v, !vsense)]: bool = isHead
				const unsigned ub = per_end[2 * (size_t)p.a + (usense ? 0 : 1)];
				const unsigned vb = per_end[2 * (size_t)p.b + (!vsense ? 0 : 1)];
				f << u << (usense ? '-' : '+') << '\t' << v << (vsense ? '-' : '+') << '\t' << (p.c[i] == max_counts ? "T" : "F") << '\t'
				  << p.c[i] << '\t' << ub << '\t' << vb << '\t' << barcodeCount << '\n';
				f << v << (vsense ? '+' : '-') << '\t' << u << (usense ? '+' : '-') << '\t' << (p.c[i] == max_counts ? "T" : "F") << '\t'
				  << p.c[i] << '\t' << vb << '\t' << ub << '\t' << barcodeCount << '\n';
			}
		}
	}

	// ---- writeBarcodeCountsTSV (Arcs.cpp:1678-1700)
	if (!params.barcode_counts_name.empty()) {
		std::cout << "\n=> Writing reads per barcode TSV file... " << stamp();
		std::string path = params.barcode_counts_name;
		if (path.find(".tsv") == std::string::npos)
			path += ".tsv";
		std::vector<std::pair<std::string, unsigned>> sorted;
		for (uint32_t i = 0; i < n_bc; ++i)
			if (bc.counted[i])
				sorted.emplace_back(bc.name[i], (unsigned)bc.mult[i]);
		std::sort(sorted.begin(), sorted.end(), [](const std::pair<std::string, unsigned>& a, const std::pair<std::string, unsigned>& b) {
			return a.second != b.second ? a.second > b.second : a.first < b.first;
		});
		std::ofstream f(path.c_str());
		if (!f)
			die("error: cannot write " + path);
		for (const auto& x : sorted)
			f << x.first << '\t' << x.second << '\n';
	}
	if (params.verbose)
		printf("wall-clock: start -> _original.gv closed %.3f s (CUDA context(s) %.3f s on a helper thread from %.3f s on, draft %.3f s%s, index %.3f s of which CUDA context + table allocation %.3f s, reads %.3f s, "
		       "pair links %.3f s, merge over %d GPUs %.3f s, export %.3f s)\n",
		    t_gv - t_start, t_ctx1 - t_ctx0, t_ctx0 - t_start, t_draft1 - t_draft0, draft_fast ? " on all cores" : " sequential reader", t_index1 - t_index0, t_gpu_init, t_map1 - t_map0, t_links1 - t_links0, params.gpus, t_merge1 - t_links1,
		    t_export1 - t_merge1);
	if (comm_init.joinable())
		comm_init.join();
	// Every output file is closed by now.  Giving 110 GB of table per GPU back to the driver piece by piece takes
	// seconds (8 GPUs: ~3 s); the operating system does the same when the process ends, so leave it to it --
	// unless ARKS_CLEAN_EXIT asks for the orderly path (sanitizer runs).
	if (!getenv("ARKS_CLEAN_EXIT")) {
		std::cout << "\n=> Done.\n" << stamp();
		std::cout.flush();
		std::cerr.flush();
		fflush(nullptr);
		_exit(0);
	}
	for (auto& gp : gpus) {
		for (auto& b : gp.batch) {
			if (!b.bases)
				continue;
			arks_host_free(b.bases);
			arks_host_free(b.off);
			arks_host_free(b.bc);
			if (b.out)
				arks_host_free(b.out);
		}
		arks_destroy(gp.h);
	}
	std::cout << "\n=> Done.\n" << stamp();
	return 0;
}
