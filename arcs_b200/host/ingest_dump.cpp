// ingest_dump.cpp -- test helper (no GPU): runs the read-pair ingest of ingest.h over one file and
// prints what it produced, so that tests/test_host_cpu.py can compare the block-parallel path with the
// faithful sequential one record for record.
//   ingest_dump <seq|par> <file> [workers] [block_bytes] [multfile.csv]
// Output: one "P <barcode> <mate1> <mate2>" line per accepted pair (input order), the counters, the
// number of blocks that went through the parallel path, and the barcode table sorted by name.
#include "ingest.h"
#include "long_cut.h"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

using namespace arks_host;

int main(int argc, char** argv)
{
	if (argc < 3)
		return 2;
	const std::string mode = argv[1]; // seq | par | seqbench | parbench (the bench modes only count)
	const bool par = mode.compare(0, 3, "par") == 0;
	const bool bench = mode.size() > 3;
	const std::string path = argv[2];
	Barcodes bc;
	IngestConfig cfg;
	cfg.min_mult = 50;
	cfg.max_mult = 10000;
	// ARKS_SHARDS=G: deal the pairs to G shards by barcode; every P line then ends with its shard
	if (const char* e = getenv("ARKS_SHARDS"))
		cfg.shards = (uint32_t)std::max(1, std::min((int)kMaxShards, atoi(e)));
	if (argc > 5) { // "barcode,count" lines: multiplicities known up front
		std::ifstream in(argv[5]);
		std::string line;
		while (getline(in, line)) {
			std::istringstream iss(line);
			std::string b, m;
			getline(iss, b, ',');
			iss >> m;
			uint32_t i = bc.intern(b);
			bc.mult[i] = atoi(m.c_str());
			bc.counted[i] = 1;
		}
		cfg.mult_known = true;
	}
	IngestCounters ctr;
	std::vector<std::string> lines;
	PairSink sink;
	size_t n_pairs = 0, n_bases = 0;
	sink.add = [&](const std::string& s1, const std::string& s2, uint32_t id) {
		n_pairs++;
		n_bases += s1.size() + s2.size();
		if (!bench)
			lines.push_back("P\t" + bc.name[id] + "\t" + s1 + "\t" + s2 +
			                (cfg.shards > 1 ? "\t" + std::to_string(barcode_shard(bc.name[id].data(), bc.name[id].size(), cfg.shards)) : ""));
	};
	sink.submit = [&](const PairBatch& b) {
		n_pairs += b.n_pairs;
		n_bases += b.n_bases;
		for (uint32_t g = 0; g < b.n_shards && !bench; ++g) {
			const uint32_t* off = b.shard_off(g);
			const uint32_t* ids = b.shard_bc(g);
			for (uint32_t i = 0; i < b.shard_pairs(g); ++i)
				lines.push_back("P\t" + bc.name[ids[i]] + "\t" + std::string(b.bases + off[2 * i], off[2 * i + 1] - off[2 * i]) + "\t" +
				                std::string(b.bases + off[2 * i + 1], off[2 * i + 2] - off[2 * i + 1]) +
				                (cfg.shards > 1 ? "\t" + std::to_string(g) : ""));
		}
	};
	bool counting = !cfg.mult_known;
	size_t fast = 0;
	// ARKS_CUT=L,M: `path` holds long reads, cut on the fly like `arcs --arks --cut L --cut_min M`
	size_t cut_l = 0, cut_m = 2000;
	if (const char* e = getenv("ARKS_CUT"))
		sscanf(e, "%zu,%zu", &cut_l, &cut_m);
	auto cut_source = [&]() -> std::unique_ptr<arks_host::ByteSource> {
		if (!cut_l)
			return nullptr;
		std::unique_ptr<arks_host::LongCutSource> src(new arks_host::LongCutSource(path, cut_l, cut_m));
		if (!src->ok())
			exit(3);
		return src;
	};
	if (par) {
		ParallelIngestOptions opt;
		opt.workers = argc > 3 ? atoi(argv[3]) : 4;
		opt.block_bytes = argc > 4 ? (size_t)atol(argv[4]) : (1u << 20);
		std::vector<std::vector<char>> mem;
		const uint32_t cap_pairs = (uint32_t)(opt.block_bytes / 16 + 16);
		for (int i = 0; i < opt.workers + 2; ++i) {
			PairBatch pb;
			mem.emplace_back(opt.block_bytes + 4096);
			pb.bases = mem.back().data();
			mem.emplace_back((2ull * cap_pairs + 1 + kMaxShards) * 4);
			pb.off = (uint32_t*)mem.back().data();
			mem.emplace_back(cap_pairs * 4ull);
			pb.bc = (uint32_t*)mem.back().data();
			pb.cap_bases = opt.block_bytes;
			pb.cap_pairs = cap_pairs;
			opt.slots.push_back(pb);
		}
		if (!ingest_parallel_blocks(path, bc, cfg, counting, ctr, sink, opt, &fast, cut_source()))
			return 3;
	} else if (cut_l) {
		SeqReader rd(cut_source(), std::string());
		ingest_sequential(rd, bc, cfg, counting, ctr, sink);
	} else {
		SeqReader rd(path);
		if (!rd.ok())
			return 3;
		ingest_sequential(rd, bc, cfg, counting, ctr, sink);
	}
	fflush(stdout);
	for (const auto& l : lines)
		puts(l.c_str());
	printf("COUNTERS unpaired=%zu emptybarcode=%zu invalidbarcode=%zu badmult=%zu count=%zu counting=%d\n", ctr.skipped_unpaired, ctr.emptybarcode,
	    ctr.invalidbarcode, ctr.skipped_badmult, ctr.count, (int)counting);
	fprintf(stderr, "FAST_BLOCKS %zu\n", fast);
	if (bench) {
		printf("pairs=%zu bases=%zu barcodes=%zu\n", n_pairs, n_bases, bc.name.size());
		return 0;
	}
	std::vector<std::string> tab;
	for (size_t i = 0; i < bc.name.size(); ++i)
		tab.push_back("B\t" + bc.name[i] + "\t" + std::to_string(bc.mult[i]) + "\t" + std::to_string((int)bc.counted[i]));
	std::sort(tab.begin(), tab.end());
	for (const auto& l : tab)
		puts(l.c_str());
	return 0;
}
