// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Driver around the reference's own hot-path functions (bcgsc/arcs 1.2.8), which
// build_ref.sh extracts verbatim by line range from /root/reference/Arcs/Arcs.{h,cpp}
// into ref_types.inc / ref_part_a.inc / ref_part_b.inc in a temp dir.  Everything in
// THIS file is ours: argument parsing for the ARKS subset of the arcs CLI, dumps of
// the intermediate containers (kmap / per-read conreci trace / imap / pmap), phase
// timing for the CPU baseline, and a Boost-free restatement of createGraph +
// boost::write_graphviz (Arcs.cpp:1475-1526,1549-1610; text format pinned by
// Examples/arks_test-demo/output/*_original.gv).  -D: the distance-estimation functions of
// Arcs/DistanceEst.h are the reference's own (ref_dist.inc); only the two that walk the Boost graph
// (addEdgeDistances :392-430, writeDistTSV :433-493) and the edge writer (Arcs.h:185-211) are
// restated here over RefGraph.
#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <getopt.h>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <limits>
#include <map>
#include <omp.h>
#include <sstream>
#include <string>
#include <time.h>
#include <tuple>
#include <unistd.h>
#include <unordered_map>
#include <utility>
#include <vector>
#include <zlib.h>

#include "Common/IOUtil.h"
#include "Common/MapUtil.h"
#include "Common/PairHash.h"
#include "Common/ReadsProcessor.h"
#include "Common/SAM.h"
#include "Common/StatUtil.h"
#include "Common/StringUtil.h"
#include "kseq.h"
#include <array>
using std::ofstream;

#include "ref_types.inc" // opens namespace ARCS, ArcsParams .. ContigToLengthIt
// google::sparse_hash_map stand-in.  The reference only ever find()s / operator[]s
// this map (Arcs.cpp:903-917,969-971), so the container is not result-bearing.
struct ContigKMap : std::unordered_map<std::string, int>
{
	void set_deleted_key(const std::string&) {}
};
} // namespace ARCS

static ARCS::ArcsParams params;

#include "ref_part_a.inc" // counters .. bestContig

// Interpose on bestContig so the per-read decision can be traced (only meaningful at -t 1).
static FILE* g_trace = NULL;
static int
bestContig_traced(ARCS::ContigKMap& kmap, std::string readseq, int k, double j, ReadsProcessor& proc)
{
	int r = bestContig(kmap, readseq, k, j, proc);
	if (g_trace)
		fprintf(g_trace, "%d\n", r);
	return r;
}
#define bestContig bestContig_traced
#include "ref_part_b.inc" // getContigKmers .. checkSignificance, TSV writers
#undef bestContig
#include "ref_dist.inc" // DistanceEst.h minus the Boost-graph functions
#include "ref_sam.inc"  // alignment mode: checkFlag .. readBAMS

// getScaffSizes (Arcs.cpp:549-568) over kseq instead of DataLayer/FastaReader: contig id -> length
static void
getScaffSizesKseq(const std::string& file, ARCS::ContigToLength& contigToLength)
{
	gzFile fp = gzopen(file.c_str(), "r");
	kseq_t* seq = kseq_init(fp);
	while (kseq_read(seq) >= 0)
		contigToLength[seq->name.s] = (int)seq->seq.l;
	kseq_destroy(seq);
	gzclose(fp);
}

// ---- Boost-free createGraph / removeDegreeNodes / write_graphviz restatement ----
struct RefEdge
{
	int u, v, orientation, weight;
	// EdgeProperties' defaults (Arcs.h:166-182)
	int minDist = std::numeric_limits<int>::min();
	int dist = std::numeric_limits<int>::max();
	int maxDist = std::numeric_limits<int>::max();
	float jaccard = -1.0f;
};
struct RefGraph
{
	std::vector<std::string> vid;
	std::vector<RefEdge> edges;
};

static void
createGraphNoBoost(const ARCS::PairMap& pmap, RefGraph& g)
{
	std::unordered_map<std::string, int> vmap;
	for (auto it = pmap.begin(); it != pmap.end(); ++it) {
		unsigned max, index;
		std::tie(max, index) = getMaxValueAndIndex(it->second);
		unsigned second = 0;
		for (unsigned i = 0; i < it->second.size(); ++i)
			if (it->second[i] != max && it->second[i] > second)
				second = it->second[i];
		if (checkSignificance(max, max + second)) {
			for (const std::string* s : { &it->first.first, &it->first.second })
				if (!vmap.count(*s)) {
					vmap[*s] = (int)g.vid.size();
					g.vid.push_back(*s);
				}
			RefEdge e;
			e.u = vmap[it->first.first];
			e.v = vmap[it->first.second];
			e.orientation = (int)index;
			e.weight = (int)max;
			g.edges.push_back(e);
		}
	}
}

// removes the vertices of degree > max_degree from g itself, as writePostRemovalGraph does (Arcs.cpp:1593-1610):
// the ABySS graph is made from what is left
static void
writeGraphNoBoost(const std::string& path, RefGraph& g, int max_degree)
{
	if (max_degree != 0) {
		std::vector<int> deg(g.vid.size(), 0);
		for (auto& e : g.edges) {
			deg[e.u]++;
			deg[e.v]++;
		}
		std::vector<int> remap(g.vid.size(), -1);
		std::vector<std::string> nv;
		for (size_t i = 0; i < g.vid.size(); ++i)
			if (deg[i] <= max_degree) {
				remap[i] = (int)nv.size();
				nv.push_back(g.vid[i]);
			}
		std::vector<RefEdge> ne;
		for (auto& e : g.edges)
			if (remap[e.u] >= 0 && remap[e.v] >= 0) {
				RefEdge n = e;
				n.u = remap[e.u];
				n.v = remap[e.v];
				ne.push_back(n);
			}
		g.vid.swap(nv);
		g.edges.swap(ne);
	}
	std::ofstream out(path.c_str());
	out << "graph G {\n";
	for (size_t i = 0; i < g.vid.size(); ++i)
		out << i << " [id=" << g.vid[i] << "];\n";
	for (auto& e : g.edges) {
		out << e.u << "--" << e.v << " [label=" << e.orientation << ", weight=" << e.weight;
		if (e.minDist != std::numeric_limits<int>::min()) // EdgePropertyWriter, Arcs.h:204-211
			out << ", d=" << e.dist << ", maxd=" << e.maxDist;
		out << "];\n";
	}
	out << "}\n";
}

// createAbyssGraph + writeAbyssGraph (Arcs.cpp:1615-1672; Graph/DotIO.h:82-114, Graph/ContigGraph.h) without Boost /
// the ABySS graph classes.  What is pinned by the reference's own code here is the VERTEX ORDER: the walk over the
// reference's `contigToLength` container (a std::unordered_map<std::string,int> filled by the reference's
// getContigKmers), two vertices per contig ("name+", "name-").  An edge u -> v also adds its complement v^ -> u^;
// out-edges are listed per vertex in insertion order.
static void
writeAbyssGraphNoBoost(const std::string& path, const ARCS::ContigToLength& contigToLength, const RefGraph& g)
{
	std::vector<std::string> vname;
	std::vector<int> vlen;
	std::unordered_map<std::string, size_t> vindex;
	for (const auto& it : contigToLength) {
		vindex[it.first] = vname.size();
		vname.push_back(it.first + "+");
		vlen.push_back(it.second);
		vname.push_back(it.first + "-");
		vlen.push_back(it.second);
	}
	struct Out
	{
		size_t to;
		int n, d;
	};
	std::vector<std::vector<Out>> adj(vname.size());
	for (const auto& e : g.edges) {
		const size_t u = vindex[g.vid[e.u]] + (e.orientation < 2 ? 1 : 0);
		const size_t v = vindex[g.vid[e.v]] + (e.orientation % 2 ? 1 : 0);
		for (const auto& o : adj[u])
			if (o.to == v) {
				std::cerr << "error: Duplicate edge: \"" << vname[u] << "\" -> \"" << vname[v] << '"' << std::endl;
				exit(EXIT_FAILURE);
			}
		int d = (int)params.gap;
		if (params.dist_est)
			d = params.dist_mode == ARCS::DIST_MEDIAN ? e.dist : e.maxDist;
		adj[u].push_back(Out{ v, e.weight, d });
		const size_t uc = u ^ 1, vc = v ^ 1;
		if (!(vc == u && uc == v))
			adj[vc].push_back(Out{ uc, e.weight, d });
	}
	std::ofstream out(path.c_str());
	out << "digraph arcs {\n";
	for (size_t i = 0; i < vname.size(); ++i)
		out << '"' << vname[i] << "\" [l=" << vlen[i] << "]\n";
	for (size_t u = 0; u < vname.size(); ++u)
		for (const auto& o : adj[u])
			out << '"' << vname[u] << "\" -> \"" << vname[o.to] << "\" [d=" << o.d << " e=" << std::fixed << std::setprecision(1)
			    << (float)params.gap << " n=" << o.n << "]\n";
	out << "}\n";
}

// addEdgeDistances (DistanceEst.h:392-430) over RefGraph
static void
addEdgeDistancesNoBoost(const PairToBarcodeStats& pairToStats, const JaccardToDist& jaccardToDist, RefGraph& g)
{
	if (jaccardToDist.empty())
		return;
	for (auto& e : g.edges) {
		auto statsIt = pairToStats.find(std::make_pair(g.vid[e.u], g.vid[e.v]));
		if (statsIt == pairToStats.end())
			continue;
		const BarcodeStats& stats = statsIt->second.at(e.orientation);
		DistanceEstimate est;
		bool success;
		std::tie(est, success) = estimateDistance(stats, jaccardToDist, params);
		if (!success)
			continue;
		e.minDist = est.minDist;
		e.dist = est.dist;
		e.maxDist = est.maxDist;
		e.jaccard = est.jaccard;
	}
}

// writeDistTSV (DistanceEst.h:433-493) over RefGraph
static void
writeDistTSVNoBoost(const std::string& path, const PairToBarcodeStats& pairToStats, const RefGraph& g)
{
	std::ofstream tsvOut(path.c_str());
	tsvOut << "contig1" << '\t' << "contig2" << '\t' << "min_dist" << '\t' << "dist" << '\t' << "max_dist" << '\t'
	       << "barcodes1" << '\t' << "barcodes2" << '\t' << "barcodes_union" << '\t' << "barcodes_intersect" << '\n';
	for (auto& e : g.edges) {
		auto pair = std::make_pair(g.vid[e.u], g.vid[e.v]);
		auto statsIt = pairToStats.find(pair);
		if (statsIt == pairToStats.end())
			continue;
		const BarcodeStats& stats = statsIt->second.at(e.orientation);
		bool sense1 = e.orientation < 2;
		bool sense2 = e.orientation % 2;
		for (int pass = 0; pass < 2; ++pass) {
			if (pass == 0)
				tsvOut << pair.first << (sense1 ? '-' : '+') << '\t' << pair.second << (sense2 ? '-' : '+') << '\t';
			else
				tsvOut << pair.second << (sense2 ? '+' : '-') << '\t' << pair.first << (sense1 ? '+' : '-') << '\t';
			if (e.jaccard >= 0)
				tsvOut << e.minDist << '\t' << e.dist << '\t' << e.maxDist << '\t';
			else
				tsvOut << "NA" << '\t' << "NA" << '\t' << "NA" << '\t';
			if (pass == 0)
				tsvOut << stats.barcodes1 << '\t' << stats.barcodes2;
			else
				tsvOut << stats.barcodes2 << '\t' << stats.barcodes1;
			tsvOut << '\t' << stats.barcodesUnion << '\t' << stats.barcodesIntersect << '\n';
		}
	}
}

static std::string
hexkey(const std::string& s)
{
	static const char* d = "0123456789abcdef";
	std::string o;
	for (unsigned char c : s) {
		o += d[c >> 4];
		o += d[c & 15];
	}
	return o;
}

static double
now()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int
main(int argc, char** argv)
{
	std::string dump_kmap, dump_trace, dump_imap, dump_pmap, timing_json, dist_gv;
	int map_repeats = 1; // --map-repeats N: run the mapping phase N times on one index (timing only when N > 1)
	static const struct option lo[] = { { "dump-kmap", required_argument, NULL, 1001 },
		                                { "dump-trace", required_argument, NULL, 1002 },
		                                { "dump-imap", required_argument, NULL, 1003 },
		                                { "dump-pmap", required_argument, NULL, 1004 },
		                                { "timing-json", required_argument, NULL, 1005 },
		                                { "arks", no_argument, NULL, 1006 },
		                                { "tsv", required_argument, NULL, 1007 },
		                                { "barcode-counts", required_argument, NULL, 1008 },
		                                { "dist_tsv", required_argument, NULL, 1009 },
		                                { "samples_tsv", required_argument, NULL, 1010 },
		                                { "dist_est", no_argument, NULL, 'D' },
		                                { "arcs", no_argument, NULL, 1011 },
		                                { "bin_size", required_argument, NULL, 'B' },
		                                { "map-repeats", required_argument, NULL, 1012 },
		                                { "dist-gv", required_argument, NULL, 1013 },
		                                { "gap", required_argument, NULL, 1014 },
		                                { NULL, 0, NULL, 0 } };
	params.arks = true;
	for (int c; (c = getopt_long(argc, argv, "f:c:l:z:b:m:d:e:r:vt:u:j:k:DB:s:", lo, NULL)) != -1;) {
		std::istringstream arg(optarg != NULL ? optarg : "");
		switch (c) {
		case 'u': arg >> params.multfile; break;
		case 'k': arg >> params.k_value; break;
		case 'j': arg >> params.j_index; break;
		case 't': arg >> params.threads; break;
		case 'f': arg >> params.file; break;
		case 'c': arg >> params.min_reads; break;
		case 'l': arg >> params.min_links; break;
		case 'z': arg >> params.min_size; break;
		case 'b': arg >> params.base_name; break;
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
		case 1001: dump_kmap = optarg; break;
		case 1002: dump_trace = optarg; break;
		case 1003: dump_imap = optarg; break;
		case 1004: dump_pmap = optarg; break;
		case 1005: timing_json = optarg; break;
		case 1006: break;
		case 1007: params.tsv_name = optarg; break;
		case 1008: params.barcode_counts_name = optarg; break;
		case 1009: params.dist_tsv = optarg; break;
		case 1010: params.dist_samples_tsv = optarg; break;
		case 'D': params.dist_est = true; break;
		case 1011: params.arks = false; break;
		case 1012: map_repeats = std::max(1, atoi(optarg)); break;
		case 1013: dist_gv = optarg; break;
		case 1014: arg >> params.gap; break;
		case 's': arg >> params.seq_id; break;
		case 'B': arg >> params.dist_bin_size; break;
		default: return 2;
		}
	}
	std::vector<std::string> filenames(argv + optind, argv + argc);
	if ((params.arks && params.file.empty()) || filenames.empty() || params.base_name.empty()) {
		fprintf(stderr, "usage: arcs_ref -f contigs.fa -b base [arcs --arks options] reads.fq[.gz]...\n");
		return 2;
	}
	omp_set_num_threads(params.threads);

	ARCS::IndexMap imap;
	ARCS::PairMap pmap;
	std::unordered_map<std::string, int> indexMultMap;
	ARCS::ContigKMap kmap;
	kmap.set_deleted_key("");
	ARCS::ContigToLength contigToLength;
	std::vector<ARCS::CI> contigRecord;

	double t0 = now(), t1 = t0, t2 = t0;
	std::vector<double> map_times;
	if (!params.arks) { // alignment mode, runArcs :1859-1871
		if (!params.file.empty())
			getScaffSizesKseq(params.file, contigToLength);
		readBAMS(filenames, imap, indexMultMap, contigToLength);
	} else {
		if (!params.multfile.empty())
			createIndexMultMap(params.multfile, indexMultMap);
		else
			readBarcodes(filenames, indexMultMap);
		t1 = now();
		contigRecord.resize(initContigArray(params.file));
		getContigKmers(params.file, kmap, contigRecord, contigToLength);
		t2 = now();
		if (!dump_trace.empty())
			g_trace = fopen(dump_trace.c_str(), "w");
		for (int rep = 0; rep < map_repeats; ++rep) {
			const double a = now();
			if (rep)
				imap.clear();
			readChroms(filenames, kmap, imap, indexMultMap, contigRecord);
			map_times.push_back(now() - a);
		}
		if (g_trace)
			fclose(g_trace);
	}
	double t3 = now();
	if (map_repeats > 1)
		t3 = t2 + map_times.back();
	pairContigs(imap, pmap, indexMultMap);
	double t4 = now();
	RefGraph g;
	createGraphNoBoost(pmap, g);
	if (params.dist_est) { // calcDistanceEstimates, Arcs.cpp:1769-1807
		DistSampleMap distSamples;
		calcDistSamples(imap, contigToLength, indexMultMap, params, distSamples);
		writeDistSamplesTSV(params.dist_samples_tsv, distSamples);
		JaccardToDist jaccardToDist;
		buildJaccardToDist(distSamples, jaccardToDist);
		PairToBarcodeStats pairToStats;
		buildPairToBarcodeStats(imap, indexMultMap, contigToLength, params, pairToStats);
		addEdgeDistancesNoBoost(pairToStats, jaccardToDist, g);
		if (!params.dist_tsv.empty())
			writeDistTSVNoBoost(params.dist_tsv, pairToStats, g);
	}
	writeGraphNoBoost(params.base_name + "_original.gv", g, params.max_degree);
	double t5 = now();
	if (!dist_gv.empty())
		writeAbyssGraphNoBoost(dist_gv, contigToLength, g);
	if (!params.tsv_name.empty()) {
		size_t barcodeCount = countBarcodes(imap, indexMultMap);
		writeTSV(params.tsv_name, imap, pmap, barcodeCount);
	}
	if (!params.barcode_counts_name.empty())
		writeBarcodeCountsTSV(params.barcode_counts_name, indexMultMap);

	if (!dump_kmap.empty()) {
		std::vector<std::pair<std::string, int>> v(kmap.begin(), kmap.end());
		std::sort(v.begin(), v.end());
		FILE* f = fopen(dump_kmap.c_str(), "w");
		for (auto& e : v)
			fprintf(f, "%s\t%d\n", hexkey(e.first).c_str(), e.second);
		fclose(f);
	}
	if (!dump_imap.empty()) {
		std::vector<std::string> lines;
		for (auto& b : imap)
			for (auto& s : b.second) {
				std::ostringstream o;
				o << b.first << '\t' << s.first.first << '\t' << (s.first.second ? 'H' : 'T') << '\t' << s.second;
				lines.push_back(o.str());
			}
		std::sort(lines.begin(), lines.end());
		FILE* f = fopen(dump_imap.c_str(), "w");
		for (auto& l : lines)
			fprintf(f, "%s\n", l.c_str());
		fclose(f);
	}
	if (!dump_pmap.empty()) {
		FILE* f = fopen(dump_pmap.c_str(), "w");
		for (auto& it : pmap)
			fprintf(f, "%s\t%s\t%u\t%u\t%u\t%u\n", it.first.first.c_str(), it.first.second.c_str(),
			        it.second[0], it.second[1], it.second[2], it.second[3]);
		fclose(f);
	}
	// machine-readable counters + phase times (CPU baseline of bench.py reads this)
	std::string runs;
	for (size_t i = 0; i < map_times.size(); ++i) {
		char buf[32];
		snprintf(buf, sizeof(buf), "%s%.6f", i ? ", " : "", map_times[i]);
		runs += buf;
	}
	FILE* tj = timing_json.empty() ? stdout : fopen(timing_json.c_str(), "w");
	fprintf(tj,
	        "{\"threads\": %u, \"t_multiplicity_s\": %.6f, \"t_index_s\": %.6f, \"t_map_s\": %.6f, "
	        "\"t_pair_s\": %.6f, \"t_graph_s\": %.6f, \"contig_kmers\": %u, \"null_kmers\": %u, "
	        "\"recorded\": %u, \"collisions\": %u, \"removed\": %u, \"unique\": %u, "
	        "\"read_kmers_valid\": %u, \"read_kmers_invalid\": %u, \"found\": %u, \"rec\": %u, "
	        "\"dups\": %u, \"pass_jaccard\": %u, \"fail_jaccard\": %u, \"pmap_size\": %zu, "
	        "\"imap_barcodes\": %zu, \"map_repeats\": %d, \"t_map_runs\": [%s]}\n",
	        params.threads, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4,
	        s_numkmersmapped + s_numkmercollisions, s_numbadkmers, s_numkmersmapped, s_numkmercollisions,
	        s_numkmersremdup, s_uniquedraftkmers, s_totalnumckmers, s_numbadckmers, s_numckmersfound,
	        s_numckmersrec, s_ckmersasdups, s_numreadspassingjaccard, s_numreadsfailjaccard, pmap.size(),
	        imap.size(), map_repeats, runs.c_str());
	if (tj != stdout)
		fclose(tj);
	return 0;
}
