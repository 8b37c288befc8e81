// ingest.h -- read-pair ingest of the `arcs --arks` drop-in: chromiumRead's record loop
// (Arcs/Arcs.cpp:1185-1265) restated twice over the same sink:
//
//  * ingest_sequential: record by record through SeqReader (the faithful reader of seq_reader.h);
//    handles everything the reference's reader handles (FASTA, multi-line records, CRLF, garbage
//    between records, truncated files, NUL bytes ...).
//  * ingest_parallel_blocks: for the shape real read files have -- strict 4-line
//    FASTQ -- a reader thread cuts the (plain or gzip) byte stream into blocks of whole read pairs,
//    worker threads parse blocks straight into pinned batch buffers (names, BX barcodes, pairing rule,
//    block-local barcode interning), and the caller's thread commits blocks in file order (global
//    barcode ids, multiplicities, submission to the GPU).  Every block is checked for strictness
//    while it is parsed; at the first block that is not strict (and for the last few lines of every
//    file) the remaining bytes go through ingest_sequential, so results never depend on which path
//    ran.  tests/test_host_cpu.py compares the two paths record for record.
#pragma once
#include "seq_reader.h"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <fcntl.h>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <unordered_map>
#include <vector>

namespace arks_host {

// ---- barcodes (ids by first appearance; indexMultMap of the reference) -------------------------
struct Barcodes
{
	// (block pipeline only: parser threads look ids up under a shared lock while the committing thread adds new
	// barcodes under an exclusive one; every other user is single-threaded and takes no lock)
	mutable std::shared_mutex mu;
	std::unordered_map<std::string, uint32_t> id;
	std::vector<std::string> name;
	std::vector<int32_t> mult;    // indexMultMap value
	std::vector<uint8_t> counted; // the barcode is a key of indexMultMap
	uint32_t intern(const std::string& b)
	{
		auto it = id.find(b);
		if (it != id.end())
			return it->second;
		uint32_t i = (uint32_t)name.size();
		id.emplace(b, i);
		name.push_back(b);
		mult.push_back(0);
		counted.push_back(0);
		return i;
	}
};

// stripReadNum (Arcs.cpp:243-254): returns the length of the name without its read number
inline size_t stripped_length(const char* name, size_t n)
{
	size_t pos = n;
	while (pos > 0 && name[pos - 1] != '/')
		--pos;
	if (pos == 0) // no '/'
		return n;
	pos -= 1; // index of the last '/'
	if (pos == 0 || pos == n - 1)
		return n;
	if (!std::isdigit((unsigned char)name[pos + 1]))
		return n;
	return pos;
}

inline void strip_read_num(std::string& name)
{
	name.resize(stripped_length(name.data(), name.size()));
}

// barcode = text after the first "BX:Z:" up to the next ' ' (Arcs.cpp:1227-1251); false if there is no tag
inline bool find_bx(const char* c, size_t n, const char** b, size_t* bn)
{
	static const char tag[] = "BX:Z:";
	const char* end = c + n;
	const char* t = std::search(c, end, tag, tag + 5);
	if (t == end)
		return false;
	const char* sp = (const char*)memchr(t, ' ', (size_t)(end - t));
	*b = t + 5;
	*bn = (size_t)((sp ? sp : end) - (t + 5));
	return true;
}

inline void extract_bx(const std::string& comment, std::string& barcode)
{
	const char* b;
	size_t bn;
	if (find_bx(comment.data(), comment.size(), &b, &bn))
		barcode.assign(b, bn);
	else
		barcode.clear();
}

// readBarcodes' counting rule for one record (Arcs.cpp:514-537)
inline void count_barcode(const SeqRecord& r, Barcodes& bc, std::string& scratch)
{
	if (r.comment.empty())
		return;
	if (r.comment.find("BX:Z:") == std::string::npos)
		return;
	extract_bx(r.comment, scratch);
	uint32_t i = bc.intern(scratch);
	bc.mult[i]++;
	bc.counted[i] = 1;
}

struct IngestCounters
{
	size_t skipped_unpaired = 0, emptybarcode = 0, invalidbarcode = 0, skipped_badmult = 0, count = 0;
};

struct IngestConfig
{
	bool mult_known = false; // multiplicities came from -u or a first pass: unknown barcodes are rejected
	bool verbose = false;
	int min_mult = 0, max_mult = 0;
	uint32_t shards = 1; // GPUs the accepted pairs are dealt to, by barcode (barcode_shard)
};

constexpr uint32_t kMaxShards = 16;

// The GPU a barcode's read pairs go to.  Barcodes are the unit of independence of chromiumRead + pairContigs
// (every result before the final sum of the pair-link maps is per barcode, Arcs.cpp:1280-1285,1384-1432), so any
// function of the barcode works; a hash of its text can be evaluated by the parser threads, which do not know
// the global barcode ids yet.
inline uint32_t barcode_shard(const char* s, size_t n, uint32_t shards)
{
	if (shards <= 1)
		return 0;
	uint64_t h = 0xCBF29CE484222325ull; // FNV-1a, then a final mix so that the low bits depend on every byte
	for (size_t i = 0; i < n; ++i)
		h = (h ^ (unsigned char)s[i]) * 0x100000001B3ull;
	h ^= h >> 32;
	h *= 0x9E3779B97F4A7C15ull;
	return (uint32_t)((h >> 33) % shards);
}

// one parsed block of read pairs in caller-provided (pinned) buffers
struct PairBatch
{
	char* bases = nullptr;
	uint32_t* off = nullptr; // 2 n_pairs + 1 offsets into bases
	uint32_t* bc = nullptr;  // barcode id per pair
	uint64_t cap_bases = 0;
	uint32_t cap_pairs = 0; // off holds 2 cap_pairs + 1 + kMaxShards entries
	uint64_t n_bases = 0;
	uint32_t n_pairs = 0;
	// with several shards the pairs are grouped by shard: shard g owns pairs [shard_pair0[g], shard_pair0[g+1]),
	// its 2 n_g + 1 offsets start at off[2 shard_pair0[g] + g]
	uint32_t n_shards = 1;
	uint32_t shard_pair0[kMaxShards + 1] = { 0 };
	const uint32_t* shard_off(uint32_t g) const { return off + 2 * (size_t)shard_pair0[g] + g; }
	const uint32_t* shard_bc(uint32_t g) const { return bc + shard_pair0[g]; }
	uint32_t shard_pairs(uint32_t g) const { return shard_pair0[g + 1] - shard_pair0[g]; }
};

struct PairSink
{
	// one accepted pair (sequential path)
	std::function<void(const std::string& s1, const std::string& s2, uint32_t barcode)> add;
	// a whole block of accepted pairs (parallel path); the buffers may be reused when it returns
	std::function<void(const PairBatch& b)> submit;
};

// ---- the faithful record loop ---------------------------------------------------------------
// `counting`: readBarcodes' multiplicity count is still running for this file (it stops at the first
// record with an empty sequence, Arcs.cpp:516/538-540)
inline void ingest_sequential(SeqReader& rd, Barcodes& bc, const IngestConfig& cfg, bool& counting, IngestCounters& ctr, PairSink& sink)
{
	SeqRecord r1, r2;
	std::string b1, b2, n1, n2, scratch;
	bool stop = false;
	while (!stop) {
		bool paired = false;
		r1.name.clear();
		r2.name.clear();
		r1.comment.clear();
		r2.comment.clear();
		int l = rd.read(r1);
		if (l >= 0) {
			r1.truncate_at_nul();
			if (counting) {
				if (l > 0)
					count_barcode(r1, bc, scratch);
				else
					counting = false;
			}
			l = rd.read(r2);
			if (l >= 0) {
				r2.truncate_at_nul();
				if (counting) {
					if (l > 0)
						count_barcode(r2, bc, scratch);
					else
						counting = false;
				}
			} else {
				r2.name.clear();
				r2.comment.clear();
				stop = true;
			}
		} else {
			stop = true;
		}
		n1 = r1.name;
		n2 = stop && l < 0 && r2.name.empty() ? std::string() : r2.name;
		strip_read_num(n1);
		strip_read_num(n2);
		if (n1 == n2) {
			paired = true;
		} else {
			std::cout << "File contains unpaired reads: " << n1 << " " << n2 << std::endl;
			ctr.skipped_unpaired++;
		}
		ctr.count += 2;
		if (cfg.verbose && ctr.count % 10000000 == 0)
			std::cout << "Processed " << ctr.count << " read pairs." << std::endl;
		if (stop)
			break;
		extract_bx(r1.comment, b1);
		extract_bx(r2.comment, b2);
		if (b1.empty() || b2.empty()) {
			ctr.emptybarcode++;
			continue;
		}
		// the reference looks barcode1 up (and counts an invalid barcode) before it asks whether the
		// reads pair and the two barcodes agree (Arcs.cpp:1253-1262)
		uint32_t id = 0;
		if (cfg.mult_known) {
			auto it = bc.id.find(b1);
			if (it == bc.id.end() || !bc.counted[it->second]) {
				ctr.invalidbarcode++;
				continue;
			}
			id = it->second;
		}
		if (!paired || b1 != b2)
			continue;
		if (cfg.mult_known) {
			const int m = bc.mult[id];
			if (!(m > cfg.min_mult || m < cfg.max_mult)) { // goodmult, Arcs.cpp:1267
				ctr.skipped_badmult++;
				continue;
			}
		} else {
			id = bc.intern(b1); // validity (is it a key of indexMultMap) is settled after the pass
		}
		sink.add(r1.seq, r2.seq, id);
	}
}

// ---- strict 4-line FASTQ blocks ------------------------------------------------------------------
struct Block
{
	uint64_t seq = 0;
	const char* data = nullptr; // whole read pairs: a multiple of 8 lines, ends with '\n'
	size_t size = 0;
	int slot = -1;
	PairBatch out; // filled by parse_block (bc = block-local ids unless cfg.mult_known)
	bool regular = false;
	std::vector<std::string> local_barcodes; // block-local barcode table ...
	std::vector<int32_t> local_counts;       // ... and the number of records carrying each (readBarcodes' count)
	std::vector<uint32_t> global_hint;       // ... and the global id of each, if the parser thread found it known already
	size_t records = 0, skipped_unpaired = 0, emptybarcode = 0, invalidbarcode = 0, skipped_badmult = 0;
	std::string messages;
};

// Parses one block.  Returns false (block.regular = false) at the first thing that is not a strict
// 4-line FASTQ record as the reference's reader would see it: '@' header, one non-empty sequence
// line that does not start with '@', '>' or '+', a '+' line, a quality line of the same length;
// no CR, no NUL.  `frozen` (only with cfg.mult_known) is read concurrently and must not change.
inline bool parse_block(Block& b, const IngestConfig& cfg, const Barcodes* frozen)
{
	b.regular = false;
	b.out.n_pairs = 0;
	b.out.n_bases = 0;
	const char* p = b.data;
	const size_t size = b.size;
	const char* const end = p + size;
	if (size == 0 || end[-1] != '\n' || memchr(p, 0, size) || memchr(p, '\r', size))
		return false;
	std::unordered_map<std::string, uint32_t> local;
	std::string last_key;
	uint32_t last_id = UINT32_MAX;
	std::string key;
	auto intern_local = [&](const char* s, size_t n) -> uint32_t {
		if (last_id != UINT32_MAX && last_key.size() == n && memcmp(last_key.data(), s, n) == 0)
			return last_id;
		key.assign(s, n);
		auto it = local.find(key);
		uint32_t id;
		if (it != local.end()) {
			id = it->second;
		} else {
			id = (uint32_t)b.local_barcodes.size();
			local.emplace(key, id);
			b.local_barcodes.push_back(key);
			b.local_counts.push_back(0);
		}
		last_key = key;
		last_id = id;
		return id;
	};
	struct Rec
	{
		const char *name, *comment, *seq;
		size_t name_n, comment_n, seq_n;
	};
	struct Pending
	{
		const char *s0, *s1;
		uint32_t n0, n1, id, shard;
	};
	std::vector<Pending> pending;
	while (p < end) {
		Rec r[2];
		for (int m = 0; m < 2; ++m) {
			if (p >= end || *p != '@')
				return false;
			const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
			if (!nl)
				return false;
			const char* h = p + 1;
			const char* q = h;
			while (q < nl && !isspace((unsigned char)*q))
				++q;
			r[m].name = h;
			r[m].name_n = (size_t)(q - h);
			r[m].comment = q < nl ? q + 1 : nl;
			r[m].comment_n = (size_t)(nl - r[m].comment);
			const char* s = nl + 1;
			const char* nl2 = s < end ? (const char*)memchr(s, '\n', (size_t)(end - s)) : nullptr;
			if (!nl2 || nl2 == s || *s == '@' || *s == '>' || *s == '+')
				return false;
			r[m].seq = s;
			r[m].seq_n = (size_t)(nl2 - s);
			const char* pl = nl2 + 1;
			if (pl >= end || *pl != '+')
				return false;
			const char* nl3 = (const char*)memchr(pl, '\n', (size_t)(end - pl));
			if (!nl3)
				return false;
			const char* ql = nl3 + 1;
			if ((size_t)(end - ql) < r[m].seq_n + 1 || ql[r[m].seq_n] != '\n')
				return false;
			if (memchr(ql, '\n', r[m].seq_n)) // quality shorter than the sequence: the reader would go on to the next line
				return false;
			p = ql + r[m].seq_n + 1;
			b.records++;
		}
		// readBarcodes' count (every record whose comment has a BX tag), then chromiumRead's pair rule
		const char* bx[2] = { nullptr, nullptr };
		size_t bxn[2] = { 0, 0 };
		uint32_t lid[2] = { 0, 0 };
		for (int m = 0; m < 2; ++m)
			if (r[m].comment_n && find_bx(r[m].comment, r[m].comment_n, &bx[m], &bxn[m])) {
				if (!cfg.mult_known) {
					lid[m] = intern_local(bx[m], bxn[m]);
					b.local_counts[lid[m]]++;
				}
			} else {
				bx[m] = nullptr;
				bxn[m] = 0;
			}
		const size_t n1 = stripped_length(r[0].name, r[0].name_n), n2 = stripped_length(r[1].name, r[1].name_n);
		const bool paired = n1 == n2 && memcmp(r[0].name, r[1].name, n1) == 0;
		if (!paired) {
			b.messages.append("File contains unpaired reads: ").append(r[0].name, n1).append(" ").append(r[1].name, n2).append("\n");
			b.skipped_unpaired++;
		}
		if (bxn[0] == 0 || bxn[1] == 0) {
			b.emptybarcode++;
			continue;
		}
		uint32_t id = 0;
		if (cfg.mult_known) { // looked up before the pair / equal-barcode test, as Arcs.cpp:1253-1262 does
			key.assign(bx[0], bxn[0]);
			auto it = frozen->id.find(key);
			if (it == frozen->id.end() || !frozen->counted[it->second]) {
				b.invalidbarcode++;
				continue;
			}
			id = it->second;
		}
		if (!paired || bxn[0] != bxn[1] || memcmp(bx[0], bx[1], bxn[0]) != 0)
			continue;
		if (cfg.mult_known) {
			const int mlt = frozen->mult[id];
			if (!(mlt > cfg.min_mult || mlt < cfg.max_mult)) {
				b.skipped_badmult++;
				continue;
			}
		} else {
			id = lid[0];
		}
		PairBatch& o = b.out;
		if (o.n_pairs >= o.cap_pairs || o.n_bases + r[0].seq_n + r[1].seq_n > o.cap_bases || o.n_bases + r[0].seq_n + r[1].seq_n > 0xFFFFFFF0ull)
			return false; // cannot happen for blocks cut by the reader; be safe
		if (cfg.shards > 1) { // laid out by shard once the block is parsed
			pending.push_back(Pending{ r[0].seq, r[1].seq, (uint32_t)r[0].seq_n, (uint32_t)r[1].seq_n, id,
			    barcode_shard(bx[0], bxn[0], cfg.shards) });
			o.n_bases += r[0].seq_n + r[1].seq_n;
			o.n_pairs++;
			continue;
		}
		o.off[2 * o.n_pairs] = (uint32_t)o.n_bases;
		memcpy(o.bases + o.n_bases, r[0].seq, r[0].seq_n);
		o.n_bases += r[0].seq_n;
		o.off[2 * o.n_pairs + 1] = (uint32_t)o.n_bases;
		memcpy(o.bases + o.n_bases, r[1].seq, r[1].seq_n);
		o.n_bases += r[1].seq_n;
		o.bc[o.n_pairs] = id;
		o.n_pairs++;
	}
	PairBatch& o = b.out;
	o.n_shards = std::max<uint32_t>(1, cfg.shards);
	if (cfg.shards > 1) {
		uint32_t pairs[kMaxShards] = { 0 };
		uint64_t bases[kMaxShards] = { 0 };
		for (const Pending& q : pending) {
			pairs[q.shard]++;
			bases[q.shard] += q.n0 + q.n1;
		}
		uint64_t base_at[kMaxShards];
		uint32_t pair_at[kMaxShards];
		uint64_t nb = 0;
		uint32_t np = 0;
		for (uint32_t g = 0; g < cfg.shards; ++g) {
			o.shard_pair0[g] = np;
			base_at[g] = nb;
			pair_at[g] = np;
			np += pairs[g];
			nb += bases[g];
		}
		o.shard_pair0[cfg.shards] = np;
		for (const Pending& q : pending) {
			const uint32_t g = q.shard;
			uint32_t* off = o.off + 2 * (size_t)pair_at[g] + g;
			off[0] = (uint32_t)base_at[g];
			memcpy(o.bases + base_at[g], q.s0, q.n0);
			base_at[g] += q.n0;
			off[1] = (uint32_t)base_at[g];
			memcpy(o.bases + base_at[g], q.s1, q.n1);
			base_at[g] += q.n1;
			off[2] = (uint32_t)base_at[g];
			o.bc[pair_at[g]] = q.id;
			pair_at[g]++;
		}
	} else {
		o.shard_pair0[0] = 0;
		o.shard_pair0[1] = o.n_pairs;
		o.off[2 * o.n_pairs] = (uint32_t)o.n_bases;
	}
	b.regular = true;
	return true;
}

// number of '\n' bytes in [p, p + n)
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
__attribute__((target("avx2"))) inline size_t count_newlines_avx2(const char* p, size_t n)
{
	const __m256i nl = _mm256_set1_epi8('\n');
	size_t c = 0, i = 0;
	for (; i + 32 <= n; i += 32)
		c += (size_t)__builtin_popcount((unsigned)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i*)(p + i)), nl)));
	for (; i < n; ++i)
		c += p[i] == '\n';
	return c;
}
#endif
inline size_t count_newlines(const char* p, size_t n)
{
#if defined(__x86_64__) && defined(__GNUC__)
	static const bool have_avx2 = __builtin_cpu_supports("avx2");
	if (have_avx2)
		return count_newlines_avx2(p, n);
#endif
	size_t c = 0;
	for (size_t i = 0; i < n; ++i)
		c += p[i] == '\n';
	return c;
}

// ---- the block pipeline --------------------------------------------------------------------------
struct ParallelIngestOptions
{
	int workers = 4;
	size_t block_bytes = 4u << 20;
	// pinned buffers for the parsed pairs, one set per block in flight (allocated by the caller)
	std::vector<PairBatch> slots;
};

// Reads a file (plain or gzip) block-wise.  start() opens it and launches the reader and the parser
// threads, which run ahead by as many blocks as there are slots (the caller can do something else in the
// meantime -- the CLI builds the k-mer index on the GPU); finish() commits the regular blocks in file order
// through sink.submit (after the block-local barcode ids have been replaced by global ones) and, from the
// first irregular block on and for the tail of the file, sends the bytes through ingest_sequential.
class ParallelIngest
{
  public:
	ParallelIngest() = default;
	ParallelIngest(const ParallelIngest&) = delete;
	ParallelIngest& operator=(const ParallelIngest&) = delete;
	~ParallelIngest()
	{
		if (m_started && !m_finished)
			abort_threads();
	}

	// `frozen` (only with cfg.mult_known): the barcode table, which must not change until finish() returns.
	// Returns false if the file cannot be opened.
	// `src`: an already opened stream to read instead of `path` (e.g. long reads cut on the fly, long_cut.h).
	bool start(const std::string& path, const IngestConfig& cfg, const ParallelIngestOptions& opt, const Barcodes* frozen,
	    std::unique_ptr<ByteSource> src = nullptr, const Barcodes* live = nullptr)
	{
		m_cfg = cfg;
		m_opt = opt;
		m_frozen = frozen;
		m_live = live;
		if (src) {
			m_src = std::move(src);
			return launch();
		}
		// plain regular files are memory-mapped (blocks are cut in place, nothing is copied); gzip files (fast
		// decoder) and pipes (zlib, transparent for plain data) are read into block buffers
		m_fd = ::open(path.c_str(), O_RDONLY);
		if (m_fd < 0)
			return false;
		unsigned char magic[2] = { 0, 0 };
		struct stat st;
		const bool regular = fstat(m_fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0;
		const bool plain = regular && pread(m_fd, magic, 2, 0) == 2 && !(magic[0] == 0x1f && magic[1] == 0x8b);
		if (plain) {
			void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, m_fd, 0);
			if (m != MAP_FAILED) {
				m_map = (const char*)m;
				m_map_size = (size_t)st.st_size;
				madvise(m, m_map_size, MADV_SEQUENTIAL);
			}
		}
		if (!m_map) {
			::close(m_fd);
			m_fd = -1;
			m_src = open_source(path);
			if (!m_src)
				return false;
		}
		return launch();
	}

  private:
	bool launch()
	{
		m_bufs.resize(m_opt.slots.size());
		for (size_t i = 0; i < m_opt.slots.size(); ++i)
			m_free_slots.push_back((int)i);
		m_started = true;
		m_reader = std::thread([this] { reader_loop(); });
		for (int w = 0; w < std::max(1, m_opt.workers); ++w)
			m_workers.emplace_back([this] { worker_loop(); });
		return true;
	}

  public:

	void finish(Barcodes& bc, bool& counting, IngestCounters& ctr, PairSink& sink, size_t* n_fast_blocks = nullptr)
	{
		// committer (this thread): blocks in file order
		uint64_t next = 0;
		size_t n_fast = 0;
		std::vector<uint32_t> gid;
		std::unique_ptr<Block> irregular;
		for (;;) {
			std::unique_ptr<Block> b;
			{
				std::unique_lock<std::mutex> lk(m_mu);
				m_cv.wait(lk, [&] { return m_parsed.count(next) || (m_reader_done && next >= m_n_cut); });
				if (!m_parsed.count(next))
					break; // everything that was cut has been committed
				b = std::move(m_parsed[next]);
				m_parsed.erase(next);
			}
			if (!b->regular) {
				// stop cutting; this block and everything behind it is re-read by the faithful reader
				{
					std::lock_guard<std::mutex> lk(m_mu);
					m_stop = true;
				}
				m_cv.notify_all();
				irregular = std::move(b);
				break;
			}
			// multiplicities and global barcode ids
			if (!m_cfg.mult_known) {
				gid.resize(b->local_barcodes.size());
				{
					// barcodes the parser thread did not find: one exclusive section per block
					bool any_unknown = false;
					for (size_t i = 0; i < gid.size(); ++i) {
						gid[i] = i < b->global_hint.size() ? b->global_hint[i] : UINT32_MAX;
						any_unknown |= gid[i] == UINT32_MAX;
					}
					if (any_unknown) {
						std::unique_lock<std::shared_mutex> lk(bc.mu);
						for (size_t i = 0; i < gid.size(); ++i)
							if (gid[i] == UINT32_MAX)
								gid[i] = bc.intern(b->local_barcodes[i]);
					}
				}
				for (size_t i = 0; i < gid.size(); ++i) {
					if (counting) {
						bc.mult[gid[i]] += b->local_counts[i];
						bc.counted[gid[i]] = 1;
					}
				}
				for (uint32_t i = 0; i < b->out.n_pairs; ++i)
					b->out.bc[i] = gid[b->out.bc[i]];
			}
			if (!b->messages.empty())
				std::cout << b->messages << std::flush;
			ctr.skipped_unpaired += b->skipped_unpaired;
			ctr.emptybarcode += b->emptybarcode;
			ctr.invalidbarcode += b->invalidbarcode;
			ctr.skipped_badmult += b->skipped_badmult;
			const size_t before = ctr.count;
			ctr.count += b->records; // the reference adds 2 per pair of records
			if (m_cfg.verbose)
				for (size_t c = (before / 10000000 + 1) * 10000000; c <= ctr.count; c += 10000000)
					std::cout << "Processed " << c << " read pairs." << std::endl;
			if (b->out.n_pairs)
				sink.submit(b->out);
			n_fast++;
			{
				std::lock_guard<std::mutex> lk(m_mu);
				m_free_slots.push_back(b->slot);
			}
			m_cv.notify_all();
			next++;
		}
		abort_threads();
		// the bytes that did not go through the block path, in file order: the irregular block, the blocks that
		// were already cut behind it, and the tail the reader holds; then whatever is still in the stream
		std::string prefix;
		if (irregular) {
			std::map<uint64_t, std::unique_ptr<Block>> rest;
			for (auto& kv : m_parsed)
				rest[kv.first] = std::move(kv.second);
			for (auto& b : m_todo)
				rest[b->seq] = std::move(b);
			prefix.assign(irregular->data, irregular->size);
			for (auto& kv : rest)
				prefix.append(kv.second->data, kv.second->size);
		}
		prefix += m_tail;
		if (n_fast_blocks)
			*n_fast_blocks = n_fast;
		if (m_map) {
			// mapped file: the stream continues behind the last block that was cut (zlib reads plain data as is)
			munmap((void*)m_map, m_map_size);
			m_map = nullptr;
			lseek(m_fd, (off_t)m_map_pos, SEEK_SET);
			m_src.reset(new ZlibSource(gzdopen(m_fd, "r")));
		}
		SeqReader rd(std::move(m_src), std::move(prefix)); // takes the stream over (and closes it)
		m_fd = -1;
		m_finished = true;
		ingest_sequential(rd, bc, m_cfg, counting, ctr, sink);
	}

  private:
	void abort_threads()
	{
		{
			std::lock_guard<std::mutex> lk(m_mu);
			m_stop = true;
		}
		m_cv.notify_all();
		if (m_reader.joinable())
			m_reader.join();
		for (auto& t : m_workers)
			if (t.joinable())
				t.join();
	}

	// cuts blocks of whole pairs = multiples of 8 lines (only meaningful for strict files; anything else is
	// caught by parse_block and sent down the sequential path)
	void reader_loop()
	{
		std::string carry;
		uint64_t seqno = 0;
		bool eof = false;
		size_t span = m_opt.block_bytes; // mapped files: bytes looked at for the next block
		while (!eof) {
			int slot;
			{
				std::unique_lock<std::mutex> lk(m_mu);
				m_cv.wait(lk, [&] { return m_stop || !m_free_slots.empty(); });
				if (m_stop)
					break;
				slot = m_free_slots.back();
				m_free_slots.pop_back();
			}
			const char* base;
			size_t have;
			if (m_map) {
				base = m_map + m_map_pos;
				have = std::min(span, m_map_size - m_map_pos);
				eof = m_map_pos + have == m_map_size;
			} else {
				std::vector<char>& buf = m_bufs[(size_t)slot];
				if (buf.size() < carry.size() + m_opt.block_bytes)
					buf.resize(carry.size() + m_opt.block_bytes);
				memcpy(buf.data(), carry.data(), carry.size());
				have = carry.size();
				const size_t want = carry.size() + m_opt.block_bytes;
				carry.clear();
				while (have < want) {
					const long n = m_src->read(buf.data() + have, want - have);
					if (n <= 0) {
						eof = true;
						break;
					}
					have += (size_t)n;
				}
				base = buf.data();
			}
			// the longest prefix made of whole groups of 8 lines: count the newlines, then step back over
			// the ones that are too many
			size_t cut = 0;
			size_t lines_in_block = 0;
			{
				size_t lines;
				if (m_map) {
					// mapped files: the newlines of a fixed grid of chunks are counted by helper threads a window
					// ahead (one thread cannot scan a file faster than ~8 GB/s); this thread only scans from the
					// last grid point in front of the block's end
					const size_t end = m_map_pos + have;
					count_ahead(end);
					const size_t g = std::min(end / kGrid, m_grid_cum.size() - 1);
					lines = (size_t)(m_grid_cum[g] + count_newlines(m_map + g * kGrid, end - g * kGrid) - m_lines_at_pos);
				} else {
					lines = count_newlines(base, have);
				}
				lines_in_block = lines >= 8 ? lines - (lines & 7u) : 0;
				size_t drop = lines & 7u; // newlines after the last whole group
				if (lines >= 8) {
					const char* last = (const char*)memrchr(base, '\n', have);
					const char* e = last + 1;
					while (drop--) {
						last = (const char*)memrchr(base, '\n', (size_t)(last - base));
						e = last + 1;
					}
					cut = (size_t)(e - base);
				}
			}
			if (m_map) {
				m_map_pos += cut;
				m_lines_at_pos += lines_in_block;
				span = cut ? m_opt.block_bytes : span + m_opt.block_bytes; // no whole pair in sight: look further
			} else {
				carry.assign(base + cut, have - cut);
			}
			if (cut == 0) {
				std::lock_guard<std::mutex> lk(m_mu);
				m_free_slots.push_back(slot);
				if (!eof && (m_map ? span : carry.size()) > 4 * m_opt.block_bytes) // lines longer than a block: not the strict shape
					m_stop = true;
				m_cv.notify_all();
				if (m_stop)
					break;
				continue;
			}
			std::unique_ptr<Block> b(new Block());
			b->seq = seqno++;
			b->data = base;
			b->size = cut;
			b->slot = slot;
			b->out = m_opt.slots[(size_t)slot];
			{
				std::lock_guard<std::mutex> lk(m_mu);
				m_todo.push_back(std::move(b));
				m_n_cut = seqno;
			}
			m_cv.notify_all();
		}
		std::lock_guard<std::mutex> lk(m_mu);
		m_tail.swap(carry);
		m_reader_done = true;
		m_cv.notify_all();
	}

	// newlines in [0, i * kGrid) for every grid point up to (at least) `upto`, counted a window at a time by a few
	// short-lived threads (which also fault the window's pages in for the parser threads)
	static constexpr size_t kGrid = 256u << 10;
	void count_ahead(size_t upto)
	{
		upto = std::min(upto, m_map_size);
		while ((m_grid_cum.size() - 1) * kGrid < upto) {
			const size_t c0 = m_grid_cum.size() - 1; // first chunk not counted yet
			const size_t total_chunks = (m_map_size + kGrid - 1) / kGrid;
			const size_t window_chunks = std::min<size_t>(total_chunks - c0, (128u << 20) / kGrid);
			const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
			const size_t nt = std::max<size_t>(1, std::min<size_t>({ 8, hw / 2, window_chunks }));
			std::vector<uint64_t> cnt(window_chunks, 0);
			auto work = [&](size_t t) {
				for (size_t c = t; c < window_chunks; c += nt) {
					const size_t a = (c0 + c) * kGrid, b = std::min(m_map_size, a + kGrid);
					cnt[c] = count_newlines(m_map + a, b - a);
				}
			};
			std::vector<std::thread> th;
			for (size_t t = 1; t < nt; ++t)
				th.emplace_back(work, t);
			work(0);
			for (auto& x : th)
				x.join();
			for (size_t c = 0; c < window_chunks; ++c)
				m_grid_cum.push_back(m_grid_cum.back() + cnt[c]);
		}
	}

	void worker_loop()
	{
		for (;;) {
			std::unique_ptr<Block> b;
			{
				std::unique_lock<std::mutex> lk(m_mu);
				m_cv.wait(lk, [&] { return !m_todo.empty() || m_reader_done || m_stop; });
				if (m_todo.empty())
					return;
				b = std::move(m_todo.front());
				m_todo.pop_front();
			}
			parse_block(*b, m_cfg, m_cfg.mult_known ? m_frozen : nullptr);
			if (b->regular && !m_cfg.mult_known && m_live && !b->local_barcodes.empty()) {
				// barcodes the run has met before are resolved here, in parallel (stLFR-style files carry thousands of
				// distinct barcodes per block: left to the committing thread alone they bound the whole ingest)
				b->global_hint.assign(b->local_barcodes.size(), UINT32_MAX);
				// (the lock is given up every few hundred look-ups so that the committing thread never waits long)
				for (size_t i0 = 0; i0 < b->local_barcodes.size(); i0 += 256) {
					std::shared_lock<std::shared_mutex> lk(m_live->mu);
					for (size_t i = i0; i < std::min(b->local_barcodes.size(), i0 + 256); ++i) {
						auto it = m_live->id.find(b->local_barcodes[i]);
						if (it != m_live->id.end())
							b->global_hint[i] = it->second;
					}
				}
			}
			{
				std::lock_guard<std::mutex> lk(m_mu);
				m_parsed[b->seq] = std::move(b);
			}
			m_cv.notify_all();
		}
	}

	IngestConfig m_cfg;
	ParallelIngestOptions m_opt;
	const Barcodes* m_frozen = nullptr;
	const Barcodes* m_live = nullptr; // the run's barcode table, read under its shared lock (one-pass mode)
	int m_fd = -1;
	std::unique_ptr<ByteSource> m_src; // gzip files and pipes
	const char* m_map = nullptr;
	size_t m_map_size = 0, m_map_pos = 0; // m_map_pos: first byte that has not been cut into a block (reader thread)
	std::vector<uint64_t> m_grid_cum = std::vector<uint64_t>(1, 0); // reader thread: see count_ahead
	uint64_t m_lines_at_pos = 0;                                    // newlines in front of m_map_pos
	std::vector<std::vector<char>> m_bufs;
	std::mutex m_mu;
	std::condition_variable m_cv;
	std::deque<std::unique_ptr<Block>> m_todo;          // cut, waiting for a worker
	std::map<uint64_t, std::unique_ptr<Block>> m_parsed; // parsed, waiting for their turn
	std::vector<int> m_free_slots;
	bool m_reader_done = false, m_stop = false, m_started = false, m_finished = false;
	std::string m_tail; // bytes after the last block the reader cut
	uint64_t m_n_cut = 0;
	std::thread m_reader;
	std::vector<std::thread> m_workers;
};

inline bool ingest_parallel_blocks(const std::string& path, Barcodes& bc, const IngestConfig& cfg, bool& counting, IngestCounters& ctr,
    PairSink& sink, ParallelIngestOptions& opt, size_t* n_fast_blocks = nullptr, std::unique_ptr<ByteSource> src = nullptr)
{
	ParallelIngest pi;
	if (!pi.start(path, cfg, opt, cfg.mult_known ? &bc : nullptr, std::move(src), cfg.mult_known ? nullptr : &bc))
		return false;
	pi.finish(bc, counting, ctr, sink, n_fast_blocks);
	return true;
}

} // namespace arks_host
