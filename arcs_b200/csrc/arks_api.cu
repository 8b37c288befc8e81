// arks_api.cu -- the extern "C" boundary of libarks_b200.so (include/arks_b200.h):
// handle, device memory, stream/event plumbing and kernel launches.  No torch types, no
// CPU fallback: every entry point either runs the sm_100a kernels or fails loudly.
#include "../../include/arks_b200.h"
#include "arks_device.cuh"
#include "arks_index.cuh"
#include "arks_links.cuh"
#include "arks_map.cuh"
#include "arks_nccl.h"
#include "arks_sort.cuh"

#include <algorithm>
#include <cctype>
#include <sched.h>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

using namespace arks;

namespace {

std::string g_create_error;

struct DevBuf
{
	void* p = nullptr;
	size_t cap = 0;
};

struct MapSlot
{
	DevBuf bases, off, bc, out;
	cudaEvent_t copied = nullptr, done = nullptr;
	bool busy = false;
	bool copy_pending = false; // arks_map_pairs_begin without its arks_map_pairs_end yet
	bool wants_out = false;
};

} // namespace

struct arks_handle
{
	int device = 0;
	int k = 0, kw = 0;
	uint64_t mask_hi = 0, mask_lo = 0;
	int sm_count = 148;
	cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
	// index
	uint8_t* table = nullptr;
	uint64_t nslots = 0;
	bool finalized = false;
	IndexCounters* d_ictr = nullptr;
	arks_index_stats istats{};
	DevBuf ib_bases, ib_off, ib_conreci, ib_inv, ib_skip, ib_tiles, ib_g0;
	// persistent packed contig text (seed-and-extend), global base coordinates
	DevBuf ct_T, ct_TINS, ct_TUNIQ, ct_end_g0, ct_end_len, ct_end_cr;
	uint64_t g_next = 0;      // next free base coordinate (multiple of 32)
	uint32_t n_global_ends = 0;
	int use_extension = 1;
	// membership prefilter over the table's keys (built by arks_index_finalize)
	unsigned long long* bloom = nullptr;
	uint64_t bloom_words = 0;
	int bloom_bits_per_key = 8;
	int lane_general = 1;
	// exact integer thresholds for the Jaccard gate and the N-fraction test
	uint32_t* d_jmin = nullptr;
	uint32_t* d_nmax = nullptr;
	double jmin_for = 0.0;
	bool jmin_valid = false;
	// map
	MapCounters* d_mctr = nullptr;
	MapSlot slots[2];
	int next_slot = 0;
	uint32_t* d_remap = nullptr;
	uint32_t n_remap = 0;
	int map_grid = 0, group_grid = 0, slow_grid = 0;
	DevBuf work; // WorkRecords of one map launch (launches on a handle are stream-ordered)
	uint32_t* d_work_count = nullptr;
	int map_mode_pair = 0;
	// imap
	unsigned long long* imap = nullptr;
	uint64_t imap_cap = 0;
	unsigned long long* d_imap_count = nullptr;
	uint64_t imap_upper = 0; // host-side upper bound on the number of rows
	// pair links: workspaces (grown on demand, kept between calls), the pmap hash, and the result as
	// device arrays sorted by (rank a, rank b): keys, counts[4], contig indices a / b
	DevBuf lk_table, lk_mult, lk_cnt, lk_fill, lk_offs, lk_sums, lk_rank, lk_inv, lk_rows, lk_rowbc;
	DevBuf st_key[2], st_val[2], st_hist, st_sums; // radix sort
	DevBuf pmap_buf;
	uint64_t pmap_cap = 0;
	unsigned long long* d_pmap_count = nullptr;
	DevBuf pm_keys, pm_counts, pm_a, pm_b;
	uint64_t pm_n = 0;
	uint32_t pm_n_contigs = 0;
	bool pmap_ready = false;
	std::vector<uint32_t> ht_table; // headOrTail decision table of the last (min_reads, error_percent)
	int ht_min_reads = -1;
	float ht_error = -1.0f;
	// multi-GPU merge (arks_merge_pmap): NCCL communicator of this handle, or null
	void* nccl_comm = nullptr;
	int comm_rank = 0, comm_size = 1;
	DevBuf mg_all, mg_dense, mg_head;
	// misc
	unsigned long long* d_scratch = nullptr; // 16 x u64 scratch counters
	std::string err;
	uint64_t launches = 0;
};

namespace {

int fail(arks_handle* h, int code, const std::string& msg)
{
	if (h)
		h->err = msg;
	else
		g_create_error = msg;
	return code;
}

#define CU(call)                                                                                          \
	do {                                                                                                  \
		cudaError_t e_ = (call);                                                                          \
		if (e_ != cudaSuccess)                                                                            \
			return fail(h, ARKS_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));              \
	} while (0)

int ensure(arks_handle* h, DevBuf& b, size_t bytes)
{
	if (b.cap >= bytes && b.p)
		return ARKS_OK;
	if (b.p) {
		CU(cudaStreamSynchronize(h->stream));
		CU(cudaFree(b.p));
		b.p = nullptr;
		b.cap = 0;
	}
	size_t want = bytes + bytes / 4 + 256;
	CU(cudaMalloc(&b.p, want));
	b.cap = want;
	return ARKS_OK;
}

// grows a persistent buffer, keeping its contents; the new tail is zero-filled
int grow_keep(arks_handle* h, DevBuf& b, size_t bytes)
{
	if (b.cap >= bytes && b.p)
		return ARKS_OK;
	size_t want = bytes + bytes / 2 + 4096;
	void* np = nullptr;
	CU(cudaMalloc(&np, want));
	CU(cudaMemsetAsync(np, 0, want, h->stream));
	if (b.p) {
		CU(cudaMemcpyAsync(np, b.p, b.cap, cudaMemcpyDeviceToDevice, h->stream));
		CU(cudaStreamSynchronize(h->stream));
		CU(cudaFree(b.p));
	}
	b.p = np;
	b.cap = want;
	return ARKS_OK;
}

ContigText contig_text(const arks_handle* h)
{
	ContigText ct;
	ct.T = (uint32_t*)h->ct_T.p;
	ct.TINS = (uint32_t*)h->ct_TINS.p;
	ct.TUNIQ = (uint32_t*)h->ct_TUNIQ.p;
	ct.end_g0 = (const uint64_t*)h->ct_end_g0.p;
	ct.end_len = (const uint32_t*)h->ct_end_len.p;
	ct.end_cr = (const uint32_t*)h->ct_end_cr.p;
	ct.n_bases = h->g_next;
	return ct;
}

int grid_for(const arks_handle* h, uint64_t work_items, int block, int per_sm)
{
	uint64_t blocks = (work_items + block - 1) / block;
	uint64_t cap = (uint64_t)h->sm_count * per_sm;
	return (int)std::max<uint64_t>(1, std::min(blocks, cap));
}

void key_masks(int k, uint64_t& mhi, uint64_t& mlo)
{
	int bits = 2 * k;
	if (bits >= 64) {
		mhi = ~0ull;
		int r = bits - 64;
		mlo = r == 0 ? 0ull : (r == 64 ? ~0ull : ~(~0ull >> r));
	} else {
		mhi = ~(~0ull >> bits);
		mlo = 0;
	}
}

// headOrTail's predicate (Arcs/Arcs.cpp:833-861) with the reference's exact types:
// mean, sd float; std::sqrt(2) double; std::erf double; result narrowed to float.
float normal_estimation(int x, float p, int n)
{
	float mean = n * p;
	float sd = std::sqrt(n * p * (1 - p));
	return 0.5 * (1 + std::erf((x - mean) / (sd * std::sqrt(2))));
}

bool ht_valid(int mx, int sum, int min_reads, float error_percent)
{
	if (sum < min_reads)
		return false;
	float cdf = normal_estimation(mx, 0.5, sum);
	return 1 - cdf < error_percent;
}

// min_max[sum] = smallest max in [ceil(sum/2), sum] that passes, UINT32_MAX if none.
// Verifies monotonicity exhaustively for sum <= 512 and around the threshold otherwise.
int build_ht_table(int min_reads, float error_percent, uint32_t n, uint32_t* out)
{
	for (uint32_t sum = 0; sum < n; ++sum) {
		int lo = (int)((sum + 1) / 2), hi = (int)sum;
		uint32_t thr = UINT32_MAX;
		if (sum <= 512) {
			for (int m = lo; m <= hi; ++m) {
				bool v = ht_valid(m, (int)sum, min_reads, error_percent);
				if (v && thr == UINT32_MAX)
					thr = (uint32_t)m;
				if (!v && thr != UINT32_MAX)
					return ARKS_E_NONMONOTONE;
			}
		} else {
			if (ht_valid(hi, (int)sum, min_reads, error_percent)) {
				int a = lo, b = hi; // invariant: b passes
				while (a < b) {
					int mid = a + (b - a) / 2;
					if (ht_valid(mid, (int)sum, min_reads, error_percent))
						b = mid;
					else
						a = mid + 1;
				}
				thr = (uint32_t)b;
				for (int m = std::max(lo, b - 16); m < b; ++m)
					if (ht_valid(m, (int)sum, min_reads, error_percent))
						return ARKS_E_NONMONOTONE;
				for (int m = b; m <= std::min(hi, b + 16); ++m)
					if (!ht_valid(m, (int)sum, min_reads, error_percent))
						return ARKS_E_NONMONOTONE;
			}
		}
		out[sum] = thr;
	}
	return ARKS_OK;
}

int imap_alloc(arks_handle* h, uint64_t cap)
{
	CU(cudaMalloc(&h->imap, cap * 16));
	// keys all-ones = empty; counters start at 0
	init_slots_kernel<<<grid_for(h, cap * 2, 256, 8), 256, 0, h->stream>>>(h->imap, cap * 2, 2);
	h->launches++;
	CU(cudaGetLastError());
	h->imap_cap = cap;
	return ARKS_OK;
}

// make room for `incoming` more rows
int imap_reserve(arks_handle* h, uint64_t incoming)
{
	if ((h->imap_upper + incoming) * 2 <= h->imap_cap) {
		h->imap_upper += incoming;
		return ARKS_OK;
	}
	unsigned long long actual = 0;
	CU(cudaMemcpyAsync(&actual, h->d_imap_count, 8, cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	h->imap_upper = actual;
	if ((h->imap_upper + incoming) * 2 > h->imap_cap) {
		uint64_t cap = h->imap_cap;
		while ((h->imap_upper + incoming) * 4 > cap)
			cap <<= 1;
		unsigned long long* old = h->imap;
		uint64_t old_cap = h->imap_cap;
		int rc = imap_alloc(h, cap);
		if (rc)
			return rc;
		CU(cudaMemsetAsync(h->d_imap_count, 0, 8, h->stream));
		imap_rehash_kernel<<<grid_for(h, old_cap, 256, 8), 256, 0, h->stream>>>(old, old_cap, h->imap, cap - 1, h->d_imap_count);
		h->launches++;
		CU(cudaGetLastError());
		CU(cudaStreamSynchronize(h->stream));
		CU(cudaFree(old));
	}
	h->imap_upper += incoming;
	return ARKS_OK;
}

// jmin[total] = smallest count with (double)count / (double)total > j  (bestContig, Arcs.cpp:997-1006)
int update_jmin(arks_handle* h, double j)
{
	if (h->jmin_valid && h->jmin_for == j)
		return ARKS_OK;
	std::vector<uint32_t> t(kRegionBases + 1, UINT32_MAX);
	for (int total = 1; total <= kRegionBases; ++total)
		for (int c = 1; c <= total; ++c)
			if ((double)c / (double)total > j) {
				t[total] = (uint32_t)c;
				break;
			}
	CU(cudaStreamSynchronize(h->stream));
	CU(cudaMemcpy(h->d_jmin, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
	h->jmin_for = j;
	h->jmin_valid = true;
	return ARKS_OK;
}

int launch_map(arks_handle* h, const char* d_bases, const uint32_t* d_off, const uint32_t* d_bc, uint32_t n_pairs, double j,
    int32_t* d_out)
{
	int rcj = update_jmin(h, j);
	if (rcj)
		return rcj;
	MapParams P{};
	P.jmin = h->d_jmin;
	P.nmax = h->d_nmax;
	P.table = h->table;
	P.nslots = h->nslots;
	P.k = (uint32_t)h->k;
	P.mask_hi = h->mask_hi;
	P.mask_lo = h->mask_lo;
	P.j_index = j;
	P.bases = d_bases;
	P.read_off = d_off;
	P.barcode_id = d_bc;
	P.n_pairs = n_pairs;
	P.conreci_out = d_out;
	P.remap = h->d_remap;
	P.n_remap = h->n_remap;
	P.imap = h->imap;
	P.imap_mask = h->imap_cap - 1;
	P.imap_count = h->d_imap_count;
	P.ctr = h->d_mctr;
	P.ct_T = (const uint32_t*)h->ct_T.p;
	P.ct_TINS = (const uint32_t*)h->ct_TINS.p;
	P.ct_TUNIQ = (const uint32_t*)h->ct_TUNIQ.p;
	P.ct_end_g0 = (const uint64_t*)h->ct_end_g0.p;
	P.ct_end_len = (const uint32_t*)h->ct_end_len.p;
	P.ct_end_cr = (const uint32_t*)h->ct_end_cr.p;
	P.ct_n_bases = h->g_next;
	P.use_extension = h->use_extension && h->ct_T.p != nullptr;
	P.bloom = h->bloom;
	P.bloom_words = h->bloom_words;
	P.lane_general = h->lane_general;
	if (h->map_mode_pair) {
		int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)n_pairs + kMapWarps - 1) / kMapWarps, (uint64_t)h->map_grid));
		if (h->kw == 1)
			map_pairs_kernel<1><<<grid, kMapThreads, 0, h->stream>>>(P);
		else
			map_pairs_kernel<2><<<grid, kMapThreads, 0, h->stream>>>(P);
	} else {
		int rcw;
		if ((rcw = ensure(h, h->work, n_pairs * sizeof(WorkRecord))))
			return rcw;
		P.work = (WorkRecord*)h->work.p;
		P.work_count = h->d_work_count;
		CU(cudaMemsetAsync(h->d_work_count, 0, 12, h->stream)); // [0] deferred pairs, [1] group tickets, [2] work-item tickets
		const uint64_t groups = ((uint64_t)n_pairs + kGroupPairs - 1) / kGroupPairs;
		int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((groups + kGroupWarps - 1) / kGroupWarps, (uint64_t)h->group_grid));
		const size_t smem = sizeof(GroupSmem) * kGroupWarps;
		int sgrid = (int)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)n_pairs + kMapWarps - 1) / kMapWarps, (uint64_t)h->slow_grid));
		if (h->kw == 1) {
			map_groups_kernel<1><<<grid, kGroupThreads, smem, h->stream>>>(P);
			map_slow_kernel<1><<<sgrid, kMapThreads, 0, h->stream>>>(P);
		} else {
			map_groups_kernel<2><<<grid, kGroupThreads, smem, h->stream>>>(P);
			map_slow_kernel<2><<<sgrid, kMapThreads, 0, h->stream>>>(P);
		}
		h->launches++;
	}
	h->launches++;
	CU(cudaGetLastError());
	return ARKS_OK;
}

int run_index_add(arks_handle* h, const char* d_bases, const uint64_t* d_end_off, const uint32_t* d_conreci,
    const uint64_t* h_end_off, uint32_t n_ends)
{
	const uint64_t n_bases = h_end_off[n_ends];
	if (n_bases == 0)
		return ARKS_OK;
	const uint64_t n_words = (n_bases + 31) / 32 + 4;
	int rc;
	if ((rc = ensure(h, h->ib_inv, n_words * 4)))
		return rc;
	if ((rc = ensure(h, h->ib_skip, n_words * 4)))
		return rc;
	// tiles of kTileWindows windows
	std::vector<IndexTile> tiles;
	for (uint32_t e = 0; e < n_ends; ++e) {
		uint64_t len = h_end_off[e + 1] - h_end_off[e];
		if (len >= (uint64_t)h->k) {
			if (len > 0xFFFFFFF0ull)
				return fail(h, ARKS_E_ARG, "contig end longer than 4 Gbp");
			uint32_t nwin = (uint32_t)(len - h->k + 1);
			for (uint32_t s = 0; s < nwin; s += kTileWindows)
				tiles.push_back(IndexTile{e, s});
		}
	}
	// global coordinates of this batch's ends in the persistent packed text (each end starts
	// at a multiple of 32 bases) and the per-end metadata
	std::vector<uint64_t> g0(n_ends);
	std::vector<uint32_t> elen(n_ends);
	uint64_t g = h->g_next;
	for (uint32_t e = 0; e < n_ends; ++e) {
		g0[e] = g;
		elen[e] = (uint32_t)(h_end_off[e + 1] - h_end_off[e]);
		g = (g + elen[e] + 31) & ~31ull;
	}
	if (g >= (1ull << kPosBits) || (uint64_t)h->n_global_ends + n_ends >= (1ull << 24))
		return fail(h, ARKS_E_ARG, "contig text too large for the slot position field");
	const uint32_t first_end = h->n_global_ends;
	if ((rc = grow_keep(h, h->ct_T, (g / 16 + 8) * 4)) || (rc = grow_keep(h, h->ct_TINS, (g / 32 + 8) * 4)) ||
	    (rc = grow_keep(h, h->ct_TUNIQ, (g / 32 + 8) * 4)) || (rc = grow_keep(h, h->ct_end_g0, ((uint64_t)first_end + n_ends) * 8)) ||
	    (rc = grow_keep(h, h->ct_end_len, ((uint64_t)first_end + n_ends) * 4)) ||
	    (rc = grow_keep(h, h->ct_end_cr, ((uint64_t)first_end + n_ends) * 4)) || (rc = ensure(h, h->ib_g0, n_ends * 8ull)))
		return rc;
	CU(cudaMemcpyAsync(h->ib_g0.p, g0.data(), n_ends * 8ull, cudaMemcpyHostToDevice, h->stream));
	CU(cudaMemcpyAsync((uint64_t*)h->ct_end_g0.p + first_end, g0.data(), n_ends * 8ull, cudaMemcpyHostToDevice, h->stream));
	CU(cudaMemcpyAsync((uint32_t*)h->ct_end_len.p + first_end, elen.data(), n_ends * 4ull, cudaMemcpyHostToDevice, h->stream));
	CU(cudaMemcpyAsync((uint32_t*)h->ct_end_cr.p + first_end, d_conreci, n_ends * 4ull, cudaMemcpyDeviceToDevice, h->stream));
	h->g_next = g;
	h->n_global_ends += n_ends;
	if (tiles.empty()) {
		CU(cudaStreamSynchronize(h->stream));
		return ARKS_OK;
	}
	if ((rc = ensure(h, h->ib_tiles, tiles.size() * sizeof(IndexTile))))
		return rc;
	CU(cudaMemcpyAsync(h->ib_tiles.p, tiles.data(), tiles.size() * sizeof(IndexTile), cudaMemcpyHostToDevice, h->stream));
	CU(cudaMemsetAsync(h->ib_skip.p, 0, n_words * 4, h->stream));
	inv_mask_kernel<<<grid_for(h, (n_bases + 31) / 32, 256, 16), 256, 0, h->stream>>>(d_bases, n_bases, (uint32_t*)h->ib_inv.p);
	walk_kernel<<<(n_ends + 127) / 128, 128, 0, h->stream>>>(d_end_off, n_ends, (uint32_t)h->k, (const uint32_t*)h->ib_inv.p,
	    (uint32_t*)h->ib_skip.p, h->d_ictr);
	int grid = (int)std::min<uint64_t>(tiles.size(), (uint64_t)h->sm_count * 8);
	if (h->kw == 1)
		insert_kernel<1><<<grid, kInsertThreads, 0, h->stream>>>((const IndexTile*)h->ib_tiles.p, (uint32_t)tiles.size(), d_bases,
		    d_end_off, d_conreci, (const uint32_t*)h->ib_skip.p, (const uint64_t*)h->ib_g0.p, first_end, contig_text(h), h->table,
		    h->nslots, (uint32_t)h->k, h->mask_hi, h->mask_lo, h->d_ictr);
	else
		insert_kernel<2><<<grid, kInsertThreads, 0, h->stream>>>((const IndexTile*)h->ib_tiles.p, (uint32_t)tiles.size(), d_bases,
		    d_end_off, d_conreci, (const uint32_t*)h->ib_skip.p, (const uint64_t*)h->ib_g0.p, first_end, contig_text(h), h->table,
		    h->nslots, (uint32_t)h->k, h->mask_hi, h->mask_lo, h->d_ictr);
	h->launches += 3;
	CU(cudaGetLastError());
	// the tile vector is pageable host memory: wait for its copy before it goes out of scope
	CU(cudaStreamSynchronize(h->stream));
	return ARKS_OK;
}

// ---- device radix sort of (key, value) records: arks_sort.cuh -------------------------------------------
// Sorts the n records in (st_key[0], st_val[0]) by the key bits [lo_bits) of each 32-bit half; returns the
// index (0/1) of the buffer pair that holds the result.
int device_sort_pairs(arks_handle* h, uint64_t n, uint32_t half_bits, int* result_buf)
{
	*result_buf = 0;
	if (n < 2)
		return ARKS_OK;
	if (n >= 0xFFFFFFFFull)
		return fail(h, ARKS_E_ARG, "pair-link map too large to sort (>= 2^32 rows)");
	const uint32_t n_tiles = (uint32_t)((n + kSortTile - 1) / kSortTile);
	const uint64_t n_hist = (uint64_t)kRadix * n_tiles;
	const uint32_t n_scan_blocks = (uint32_t)((n_hist + kScanBlock - 1) / kScanBlock);
	int rc;
	if ((rc = ensure(h, h->st_key[1], n * 8)) || (rc = ensure(h, h->st_val[1], n * 4)) || (rc = ensure(h, h->st_hist, (n_hist + 1) * 4)) ||
	    (rc = ensure(h, h->st_sums, (uint64_t)n_scan_blocks * 4)))
		return rc;
	int cur = 0;
	for (int half = 0; half < 2; ++half)
		for (uint32_t s = 0; s < half_bits; s += 8) {
			const uint32_t shift = 32u * half + s;
			auto* kin = (unsigned long long*)h->st_key[cur].p;
			auto* kout = (unsigned long long*)h->st_key[cur ^ 1].p;
			auto* vin = (uint32_t*)h->st_val[cur].p;
			auto* vout = (uint32_t*)h->st_val[cur ^ 1].p;
			auto* hist = (uint32_t*)h->st_hist.p;
			auto* sums = (uint32_t*)h->st_sums.p;
			radix_hist_kernel<<<n_tiles, kSortThreads, 0, h->stream>>>(kin, n, shift, hist, n_tiles);
			scan_block_sums_kernel<<<n_scan_blocks, kScanBlock, 0, h->stream>>>(hist, n_hist, sums);
			scan_sums_kernel<<<1, kScanBlock, 0, h->stream>>>(sums, n_scan_blocks);
			scan_apply_kernel<<<n_scan_blocks, kScanBlock, 0, h->stream>>>(hist, n_hist, sums, hist);
			radix_scatter_kernel<<<n_tiles, kSortThreads, 0, h->stream>>>(kin, vin, kout, vout, n, shift, hist, n_tiles);
			h->launches += 5;
			cur ^= 1;
		}
	CU(cudaGetLastError());
	*result_buf = cur;
	return ARKS_OK;
}

uint32_t bits_for(uint32_t n)
{
	uint32_t b = 1;
	while (b < 32 && (1ull << b) < n)
		++b;
	return b;
}

void comm_release(arks_handle* h)
{
	if (h->nccl_comm && nccl_api().ok)
		nccl_api().CommDestroy((ncclComm_t)h->nccl_comm);
	h->nccl_comm = nullptr;
	h->comm_size = 1;
	h->comm_rank = 0;
}

// contig indices of the sorted keys (rank -> first contig of that name)
int pmap_fill_names(arks_handle* h)
{
	int rc;
	if ((rc = ensure(h, h->pm_a, std::max<uint64_t>(h->pm_n, 1) * 4)) || (rc = ensure(h, h->pm_b, std::max<uint64_t>(h->pm_n, 1) * 4)))
		return rc;
	if (h->pm_n) {
		pmap_names_kernel<<<grid_for(h, h->pm_n, 256, 8), 256, 0, h->stream>>>((const unsigned long long*)h->pm_keys.p, h->pm_n,
		    (const uint32_t*)h->lk_inv.p, (uint32_t*)h->pm_a.p, (uint32_t*)h->pm_b.p);
		h->launches++;
		CU(cudaGetLastError());
	}
	return ARKS_OK;
}

} // namespace

extern "C" {

int arks_create(int device, int k, uint64_t max_kmers, arks_handle** out)
{
	arks_handle* h = nullptr;
	if (!out)
		return fail(nullptr, ARKS_E_ARG, "out is null");
	*out = nullptr;
	if (k < ARKS_MIN_K || k > ARKS_MAX_K)
		return fail(nullptr, ARKS_E_ARG, "k must be in [4, 64]");
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0)
		return fail(nullptr, ARKS_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)");
	if (device < 0 || device >= ndev)
		return fail(nullptr, ARKS_E_ARG, "bad device ordinal");
	h = new arks_handle();
	auto bail = [&](int code) {
		std::string msg = h->err;
		arks_destroy(h);
		g_create_error = msg;
		return code;
	};
#define CUC(call)                                                                        \
	do {                                                                                 \
		cudaError_t e_ = (call);                                                         \
		if (e_ != cudaSuccess) {                                                         \
			h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                 \
			return bail(ARKS_E_CUDA);                                                    \
		}                                                                                \
	} while (0)
	h->device = device;
	h->k = k;
	h->kw = k <= 32 ? 1 : 2;
	key_masks(k, h->mask_hi, h->mask_lo);
	// ARKS_TIMING=1: where the start-up time goes (context creation, table allocation, ...)
	const bool timing = getenv("ARKS_TIMING") != nullptr;
	auto tnow = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double t_last = tnow();
	auto lap = [&](const char* what) {
		if (timing) {
			const double t = tnow();
			fprintf(stderr, "arks_create: %-28s %.3f s\n", what, t - t_last);
			t_last = t;
		}
	};
	CUC(cudaSetDevice(device));
	CUC(cudaFree(0));
	lap("context");
	cudaDeviceProp prop;
	CUC(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10) {
		h->err = "device is not sm_100 (Blackwell); this library contains sm_100a code only";
		return bail(ARKS_E_CUDA);
	}
	h->sm_count = prop.multiProcessorCount;
	if (const char* s = getenv("ARKS_STACK"))
		CUC(cudaDeviceSetLimit(cudaLimitStackSize, (size_t)atoi(s)));
	CUC(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
	CUC(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
	h->stream = h->own_stream;
	for (auto& s : h->slots) {
		CUC(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
		CUC(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
	}
	// slots per key.  Every step of a probe sequence is a dependent memory round trip for one lane of a warp, so the
	// table is kept as sparse as the device's memory comfortably allows (configs[1], 12.5 M pairs: load 0.7 1.62,
	// 0.5 1.78, 0.35 1.85, 0.25 1.88, 0.125 1.92 x 1e11 k-mers/s): 0.125 if that takes at most a quarter of the free
	// memory, else 0.25 (45 %), else 0.35 (60 %), else 0.5 (80 %), else as dense as it has to be, up to 0.85 -- a 3 Gbp draft has 3e9 keys, 96 GB of slots
	// at load 1 (absent keys rarely reach the table: the membership filter answers first)
	double load = 0.5;
	{
		size_t free_b = 0, total_b = 0;
		CUC(cudaMemGetInfo(&free_b, &total_b));
		const double dense = (double)max_kmers * kSlotBytes; // bytes at load 1
		if (dense / 0.125 <= 0.25 * (double)free_b)
			load = 0.125;
		else if (dense / 0.25 <= 0.45 * (double)free_b)
			load = 0.25;
		else if (dense / 0.35 <= 0.60 * (double)free_b)
			load = 0.35;
		else if (dense / 0.5 <= 0.80 * (double)free_b)
			load = 0.5;
		else // the rest (20 %): filter, packed text, build scratch, read batches, tallies
			load = std::min(0.85, dense / (0.80 * (double)free_b));
	}
	if (const char* s = getenv("ARKS_TABLE_LOAD")) {
		double v = atof(s);
		if (v > 0.05 && v < 0.95)
			load = v;
	}
	h->nslots = std::max<uint64_t>(1024, (uint64_t)((double)max_kmers / load) + 1);
	const size_t slot_bytes = kSlotBytes;
	if (const char* s = getenv("ARKS_NO_EXTEND"))
		h->use_extension = atoi(s) ? 0 : 1;
	if (const char* s = getenv("ARKS_BLOOM_BITS"))
		h->bloom_bits_per_key = std::max(0, std::min(64, atoi(s)));
	if (const char* s = getenv("ARKS_LANE_GENERAL"))
		h->lane_general = atoi(s) ? 1 : 0;
	lap("streams, events, limits");
	CUC(cudaMalloc(&h->table, h->nslots * slot_bytes));
	lap("table cudaMalloc");
	CUC(cudaMemsetAsync(h->table, 0xFF, h->nslots * slot_bytes, h->stream));
	CUC(cudaMalloc(&h->d_ictr, sizeof(IndexCounters)));
	CUC(cudaMemsetAsync(h->d_ictr, 0, sizeof(IndexCounters), h->stream));
	CUC(cudaMalloc(&h->d_mctr, sizeof(MapCounters)));
	CUC(cudaMemsetAsync(h->d_mctr, 0, sizeof(MapCounters), h->stream));
	CUC(cudaMalloc(&h->d_imap_count, 8));
	CUC(cudaMemsetAsync(h->d_imap_count, 0, 8, h->stream));
	CUC(cudaMalloc(&h->d_pmap_count, 8));
	CUC(cudaMalloc(&h->d_scratch, 128));
	if (imap_alloc(h, 1ull << 16))
		return bail(ARKS_E_CUDA);
	{
		// nmax[len] = largest N count with !((double)n / (double)len > 0.02)  (checkReadSequence, Arcs.cpp:383-386)
		std::vector<uint32_t> t(kRegionBases + 1, 0);
		for (int len = 1; len <= kRegionBases; ++len)
			for (int n = 0; n <= len; ++n) {
				double ar = (double)n / (double)len;
				if (ar > 0.02)
					break;
				t[len] = (uint32_t)n;
			}
		CUC(cudaMalloc(&h->d_jmin, (kRegionBases + 1) * 4));
		CUC(cudaMalloc(&h->d_nmax, (kRegionBases + 1) * 4));
		CUC(cudaMemcpy(h->d_nmax, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
	}
	int per_sm = 0;
	if (h->kw == 1)
		CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, map_pairs_kernel<1>, kMapThreads, 0));
	else
		CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, map_pairs_kernel<2>, kMapThreads, 0));
	h->map_grid = h->sm_count * std::max(1, per_sm);
	{
		const size_t smem = sizeof(GroupSmem) * kGroupWarps;
		int per_sm_g = 0;
		if (h->kw == 1) {
			CUC(cudaFuncSetAttribute(map_groups_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
			CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_g, map_groups_kernel<1>, kGroupThreads, smem));
		} else {
			CUC(cudaFuncSetAttribute(map_groups_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
			CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_g, map_groups_kernel<2>, kGroupThreads, smem));
		}
		h->group_grid = h->sm_count * std::max(1, per_sm_g);
		int per_sm_s = 0;
		if (h->kw == 1)
			CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_s, map_slow_kernel<1>, kMapThreads, 0));
		else
			CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_s, map_slow_kernel<2>, kMapThreads, 0));
		h->slow_grid = h->sm_count * std::max(1, per_sm_s);
		CUC(cudaMalloc(&h->d_work_count, 16));
		if (const char* s = getenv("ARKS_MAP_MODE"))
			h->map_mode_pair = strcmp(s, "pair") == 0;
	}
	lap("small buffers, occupancy");
	CUC(cudaStreamSynchronize(h->stream));
	lap("table fill");
#undef CUC
	*out = h;
	return ARKS_OK;
}

void arks_destroy(arks_handle* h)
{
	if (!h)
		return;
	cudaSetDevice(h->device);
	if (h->stream)
		cudaStreamSynchronize(h->stream);
	if (h->copy_stream)
		cudaStreamSynchronize(h->copy_stream);
	for (DevBuf* b : {&h->ib_bases, &h->ib_off, &h->ib_conreci, &h->ib_inv, &h->ib_skip, &h->ib_tiles, &h->ib_g0, &h->ct_T, &h->ct_TINS,
	         &h->ct_TUNIQ, &h->ct_end_g0, &h->ct_end_len, &h->ct_end_cr})
		if (b->p)
			cudaFree(b->p);
	for (auto& s : h->slots) {
		for (DevBuf* b : {&s.bases, &s.off, &s.bc, &s.out})
			if (b->p)
				cudaFree(b->p);
		if (s.copied)
			cudaEventDestroy(s.copied);
		if (s.done)
			cudaEventDestroy(s.done);
	}
	for (DevBuf* b : {&h->lk_table, &h->lk_mult, &h->lk_cnt, &h->lk_fill, &h->lk_offs, &h->lk_sums, &h->lk_rank, &h->lk_inv, &h->lk_rows,
	         &h->lk_rowbc, &h->st_key[0], &h->st_key[1], &h->st_val[0], &h->st_val[1], &h->st_hist, &h->st_sums, &h->pmap_buf, &h->pm_keys,
	         &h->pm_counts, &h->pm_a, &h->pm_b, &h->mg_all, &h->mg_dense, &h->mg_head})
		if (b->p)
			cudaFree(b->p);
	comm_release(h);
	for (void* p : {(void*)h->table, (void*)h->d_ictr, (void*)h->d_mctr, (void*)h->d_remap, (void*)h->imap, (void*)h->d_imap_count,
	         (void*)h->d_pmap_count, (void*)h->d_scratch, (void*)h->d_jmin, (void*)h->d_nmax, (void*)h->d_work_count,
	         h->work.p, (void*)h->bloom})
		if (p)
			cudaFree(p);
	if (h->own_stream)
		cudaStreamDestroy(h->own_stream);
	if (h->copy_stream)
		cudaStreamDestroy(h->copy_stream);
	delete h;
}

const char* arks_last_error(const arks_handle* h)
{
	return h ? h->err.c_str() : g_create_error.c_str();
}

int arks_set_stream(arks_handle* h, void* cuda_stream)
{
	if (!h)
		return ARKS_E_ARG;
	CU(cudaSetDevice(h->device));
	CU(cudaStreamSynchronize(h->stream));
	h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
	return ARKS_OK;
}

int arks_sync(arks_handle* h)
{
	if (!h)
		return ARKS_E_ARG;
	CU(cudaSetDevice(h->device));
	CU(cudaStreamSynchronize(h->copy_stream));
	CU(cudaStreamSynchronize(h->stream));
	return ARKS_OK;
}

int arks_host_alloc(void** p, size_t bytes)
{
	arks_handle* h = nullptr;
	if (!p)
		return ARKS_E_ARG;
	// portable: pinned for every device of the process (one parsed block feeds several GPUs in `arcs --gpus N`)
	CU(cudaHostAlloc(p, bytes, cudaHostAllocPortable));
	return ARKS_OK;
}

int arks_device_init(int device)
{
	arks_handle* h = nullptr;
	CU(cudaSetDevice(device));
	CU(cudaFree(0));
	return ARKS_OK;
}

// NUMA node of a CUDA device from sysfs (-1 if it cannot be told)
static int device_numa_node(int device)
{
	char bus[32] = {0};
	if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess)
		return -1;
	for (char* c = bus; *c; ++c)
		*c = (char)tolower((unsigned char)*c);
	const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
	FILE* f = fopen(path.c_str(), "r");
	if (!f)
		return -1;
	int node = -1;
	if (fscanf(f, "%d", &node) != 1)
		node = -1;
	fclose(f);
	return node;
}

int arks_device_numa_node(int device, int* node)
{
	if (!node)
		return ARKS_E_ARG;
	*node = device_numa_node(device);
	return ARKS_OK;
}

int arks_bind_thread(int device)
{
	const int node = device_numa_node(device);
	if (node < 0)
		return ARKS_OK; // single-node host or unknown topology: nothing to do
	char path[128];
	snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
	FILE* f = fopen(path, "r");
	if (!f)
		return ARKS_OK;
	char list[4096] = {0};
	const bool got = fgets(list, sizeof(list), f) != nullptr;
	fclose(f);
	if (!got)
		return ARKS_OK;
	cpu_set_t set;
	CPU_ZERO(&set);
	int n_set = 0;
	for (char* tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
		int a = 0, b = 0;
		const int got2 = sscanf(tok, "%d-%d", &a, &b);
		if (got2 == 1)
			b = a;
		if (got2 >= 1)
			for (int c = a; c <= b && c < CPU_SETSIZE; ++c) {
				CPU_SET(c, &set);
				n_set++;
			}
	}
	if (n_set)
		sched_setaffinity(0, sizeof(set), &set); // the calling thread; best effort (a cgroup may forbid some cores)
	return ARKS_OK;
}

int arks_host_free(void* p)
{
	arks_handle* h = nullptr;
	CU(cudaFreeHost(p));
	return ARKS_OK;
}

// debugging aid (not in the public header): copies internal scratch to the host
int arks_debug_copy(arks_handle* h, int what, void* dst, size_t bytes)
{
	const void* src = what == 0 ? (const void*)h->d_work_count : h->work.p;
	CU(cudaStreamSynchronize(h->stream));
	CU(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
	return ARKS_OK;
}

uint64_t arks_launch_count(const arks_handle* h)
{
	return h ? h->launches : 0;
}

// ---- kernel 1 ----------------------------------------------------------------------------

int arks_index_add(arks_handle* h, const char* bases, const uint64_t* end_off, const uint32_t* conreci, uint32_t n_ends)
{
	if (!h || !bases || !end_off || !conreci)
		return fail(h, ARKS_E_ARG, "arks_index_add: null argument");
	if (h->finalized)
		return fail(h, ARKS_E_STATE, "arks_index_add after arks_index_finalize");
	if (n_ends == 0)
		return ARKS_OK;
	CU(cudaSetDevice(h->device));
	for (uint32_t e = 0; e < n_ends; ++e)
		if (conreci[e] == 0 || conreci[e] >= 0x7FFFFFFFu || end_off[e + 1] < end_off[e])
			return fail(h, ARKS_E_ARG, "arks_index_add: conreci must be >= 1 and offsets non-decreasing");
	const uint64_t n_bases = end_off[n_ends] - end_off[0];
	int rc;
	if ((rc = ensure(h, h->ib_bases, n_bases + 64)) || (rc = ensure(h, h->ib_off, (n_ends + 1) * 8ull)) ||
	    (rc = ensure(h, h->ib_conreci, n_ends * 4ull)))
		return rc;
	std::vector<uint64_t> rel(n_ends + 1);
	for (uint32_t e = 0; e <= n_ends; ++e)
		rel[e] = end_off[e] - end_off[0];
	CU(cudaMemcpyAsync(h->ib_bases.p, bases + end_off[0], n_bases, cudaMemcpyHostToDevice, h->stream));
	CU(cudaMemcpyAsync(h->ib_off.p, rel.data(), (n_ends + 1) * 8ull, cudaMemcpyHostToDevice, h->stream));
	CU(cudaMemcpyAsync(h->ib_conreci.p, conreci, n_ends * 4ull, cudaMemcpyHostToDevice, h->stream));
	rc = run_index_add(h, (const char*)h->ib_bases.p, (const uint64_t*)h->ib_off.p, (const uint32_t*)h->ib_conreci.p, rel.data(), n_ends);
	if (rc)
		return rc;
	CU(cudaStreamSynchronize(h->stream));
	return ARKS_OK;
}

int arks_index_add_device(arks_handle* h, const char* d_bases, const uint64_t* d_end_off, const uint32_t* d_conreci,
    const uint64_t* h_end_off, uint32_t n_ends)
{
	if (!h || !d_bases || !d_end_off || !d_conreci || !h_end_off)
		return fail(h, ARKS_E_ARG, "arks_index_add_device: null argument");
	if (h->finalized)
		return fail(h, ARKS_E_STATE, "arks_index_add after arks_index_finalize");
	if (n_ends == 0)
		return ARKS_OK;
	if (h_end_off[0] != 0)
		return fail(h, ARKS_E_ARG, "arks_index_add_device: end_off[0] must be 0");
	CU(cudaSetDevice(h->device));
	return run_index_add(h, d_bases, d_end_off, d_conreci, h_end_off, n_ends);
}

int arks_index_finalize(arks_handle* h, arks_index_stats* stats)
{
	if (!h)
		return ARKS_E_ARG;
	CU(cudaSetDevice(h->device));
	if (!h->finalized) {
		int grid = grid_for(h, h->nslots, 256, 16);
		// the membership filter is sized from the number of k-mers inserted (>= the number of distinct keys, and
		// within a per cent of it for any real draft), so that it can be filled by the same pass that freezes the table
		IndexCounters c;
		CU(cudaMemcpyAsync(&c, h->d_ictr, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
		CU(cudaStreamSynchronize(h->stream));
		if (c.probe_fail)
			return fail(h, ARKS_E_CAPACITY, "index table full: more distinct k-mers than max_kmers allows");
		if (h->bloom_bits_per_key > 0 && c.kmers_valid > 0) {
			// a dense table (huge drafts: load 0.5 and above) makes every false positive of the filter a walk over
			// occupied slots: twice the bits per key then -- the filter is far from L2-resident at that size anyway
			// (2 Gbp draft, load 0.5: 1.59 -> 1.65e11 k-mers/s; with a sparse table the wider filter only costs
			// locality: c5, load 0.25: 3.90 -> 3.84e10)
			int bits = h->bloom_bits_per_key;
			if (!getenv("ARKS_BLOOM_BITS") && (double)c.kmers_valid > 0.45 * (double)h->nslots)
				bits *= 2;
			h->bloom_words = std::max<uint64_t>(1024, (c.kmers_valid * (uint64_t)bits + 63) / 64);
			CU(cudaMalloc(&h->bloom, h->bloom_words * 8));
			CU(cudaMemsetAsync(h->bloom, 0, h->bloom_words * 8, h->stream));
		}
		uint32_t* tuniq = (h->g_next && h->ct_T.p) ? (uint32_t*)h->ct_TUNIQ.p : nullptr;
		if (h->kw == 1)
			finalize_kernel<1><<<grid, 256, 0, h->stream>>>(h->table, h->nslots, h->d_ictr, h->bloom, h->bloom_words, tuniq);
		else
			finalize_kernel<2><<<grid, 256, 0, h->stream>>>(h->table, h->nslots, h->d_ictr, h->bloom, h->bloom_words, tuniq);
		h->launches++;
		CU(cudaGetLastError());
		if (tuniq) {
			int g2 = grid_for(h, (h->g_next + 31) / 32 * 32, 256, 8);
			if (h->kw == 1)
				uniq_mask_kernel<1><<<g2, 256, 0, h->stream>>>(contig_text(h), h->table, h->nslots, (uint32_t)h->k, h->mask_hi, h->mask_lo);
			else
				uniq_mask_kernel<2><<<g2, 256, 0, h->stream>>>(contig_text(h), h->table, h->nslots, (uint32_t)h->k, h->mask_hi, h->mask_lo);
			h->launches++;
			CU(cudaGetLastError());
		}
		CU(cudaMemcpyAsync(&c, h->d_ictr, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
		CU(cudaStreamSynchronize(h->stream));
		h->istats.kmers_valid = c.kmers_valid;
		h->istats.kmers_null = c.kmers_null;
		h->istats.recorded = c.recorded;
		h->istats.collisions = c.kmers_valid - c.recorded;
		h->istats.removed = c.kmers_valid - c.sum_cmin;
		h->istats.unique = c.unique;
		h->finalized = true;
	}
	if (stats)
		*stats = h->istats;
	return ARKS_OK;
}

int arks_index_size(arks_handle* h, uint64_t* n_keys)
{
	if (!h || !n_keys)
		return ARKS_E_ARG;
	if (!h->finalized)
		return fail(h, ARKS_E_STATE, "arks_index_size before arks_index_finalize");
	*n_keys = h->istats.recorded;
	return ARKS_OK;
}

int arks_index_dump(arks_handle* h, uint8_t* keys, int32_t* values, uint64_t cap, uint64_t* n_keys)
{
	if (!h || !keys || !values || !n_keys)
		return ARKS_E_ARG;
	if (!h->finalized)
		return fail(h, ARKS_E_STATE, "arks_index_dump before arks_index_finalize");
	CU(cudaSetDevice(h->device));
	const uint64_t n = h->istats.recorded;
	*n_keys = n;
	if (cap < n)
		return fail(h, ARKS_E_ARG, "arks_index_dump: cap too small");
	if (n == 0)
		return ARKS_OK;
	uint64_t *d_hi = nullptr, *d_lo = nullptr;
	int32_t* d_v = nullptr;
	CU(cudaMalloc(&d_hi, n * 8));
	CU(cudaMalloc(&d_lo, n * 8));
	CU(cudaMalloc(&d_v, n * 4));
	CU(cudaMemsetAsync(h->d_scratch, 0, 8, h->stream));
	int grid = grid_for(h, h->nslots, 256, 16);
	if (h->kw == 1)
		dump_kernel<1><<<grid, 256, 0, h->stream>>>(h->table, h->nslots, d_hi, d_lo, d_v, h->d_scratch, n);
	else
		dump_kernel<2><<<grid, 256, 0, h->stream>>>(h->table, h->nslots, d_hi, d_lo, d_v, h->d_scratch, n);
	h->launches++;
	CU(cudaGetLastError());
	std::vector<uint64_t> hi(n), lo(n);
	CU(cudaMemcpyAsync(hi.data(), d_hi, n * 8, cudaMemcpyDeviceToHost, h->stream));
	CU(cudaMemcpyAsync(lo.data(), d_lo, n * 8, cudaMemcpyDeviceToHost, h->stream));
	CU(cudaMemcpyAsync(values, d_v, n * 4, cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	cudaFree(d_hi);
	cudaFree(d_lo);
	cudaFree(d_v);
	const int nb = (h->k + 3) / 4;
	for (uint64_t i = 0; i < n; ++i)
		for (int b = 0; b < nb; ++b)
			keys[i * nb + b] = (uint8_t)(b < 8 ? hi[i] >> (56 - 8 * b) : lo[i] >> (56 - 8 * (b - 8)));
	return ARKS_OK;
}

// ---- kernel 2 ----------------------------------------------------------------------------

int arks_set_conreci_remap(arks_handle* h, const uint32_t* remap, uint32_t n)
{
	if (!h)
		return ARKS_E_ARG;
	CU(cudaSetDevice(h->device));
	CU(cudaStreamSynchronize(h->stream));
	if (h->d_remap) {
		CU(cudaFree(h->d_remap));
		h->d_remap = nullptr;
		h->n_remap = 0;
	}
	if (remap && n) {
		for (uint32_t c = 1; c < n; ++c)
			if (remap[c] == 0 || ((remap[c] ^ c) & 1u))
				return fail(h, ARKS_E_ARG, "remap must keep head/tail parity and be >= 1");
		CU(cudaMalloc(&h->d_remap, n * 4ull));
		CU(cudaMemcpy(h->d_remap, remap, n * 4ull, cudaMemcpyHostToDevice));
		h->n_remap = n;
	}
	return ARKS_OK;
}

int arks_map_pairs(arks_handle* h, const char* bases, const uint32_t* read_off, const uint32_t* barcode_id, uint32_t n_pairs,
    double j_index, int32_t* conreci_out)
{
	int rc = arks_map_pairs_begin(h, bases, read_off, barcode_id, n_pairs, j_index, conreci_out);
	if (rc)
		return rc;
	return arks_map_pairs_end(h);
}

int arks_map_pairs_end(arks_handle* h)
{
	if (!h)
		return ARKS_E_ARG;
	CU(cudaSetDevice(h->device));
	for (auto& s : h->slots)
		if (s.copy_pending) {
			// inputs consumed once the copies have landed; the kernels keep running
			CU(cudaEventSynchronize(s.copied));
			s.copy_pending = false;
			if (s.wants_out) {
				CU(cudaStreamSynchronize(h->stream));
				s.wants_out = false;
			}
		}
	return ARKS_OK;
}

int arks_map_pairs_begin(arks_handle* h, const char* bases, const uint32_t* read_off, const uint32_t* barcode_id, uint32_t n_pairs,
    double j_index, int32_t* conreci_out)
{
	if (!h || !bases || !read_off || !barcode_id)
		return fail(h, ARKS_E_ARG, "arks_map_pairs: null argument");
	if (!h->finalized)
		return fail(h, ARKS_E_STATE, "arks_map_pairs before arks_index_finalize");
	if (n_pairs == 0)
		return ARKS_OK;
	if (n_pairs > 0x7FFFFFFFu)
		return fail(h, ARKS_E_ARG, "arks_map_pairs: at most 2^31-1 pairs per call");
	CU(cudaSetDevice(h->device));
	const uint32_t first = read_off[0];
	const uint64_t n_bases = read_off[2ull * n_pairs] - first;
	MapSlot& s = h->slots[h->next_slot];
	h->next_slot ^= 1;
	if (s.copy_pending) { // a begin without its end: its buffers are the caller's business, the slot's are ours
		CU(cudaEventSynchronize(s.copied));
		s.copy_pending = false;
	}
	if (s.busy) { // the kernel that last read this slot's buffers must be finished
		CU(cudaEventSynchronize(s.done));
		s.busy = false;
	}
	int rc;
	if ((rc = ensure(h, s.bases, n_bases + 64)) || (rc = ensure(h, s.off, (2ull * n_pairs + 1) * 4)) ||
	    (rc = ensure(h, s.bc, n_pairs * 4ull)) || (conreci_out && (rc = ensure(h, s.out, n_pairs * 4ull))))
		return rc;
	if ((rc = imap_reserve(h, n_pairs)))
		return rc;
	CU(cudaMemcpyAsync(s.bases.p, bases + first, n_bases, cudaMemcpyHostToDevice, h->copy_stream));
	CU(cudaMemcpyAsync(s.off.p, read_off, (2ull * n_pairs + 1) * 4, cudaMemcpyHostToDevice, h->copy_stream));
	CU(cudaMemcpyAsync(s.bc.p, barcode_id, n_pairs * 4ull, cudaMemcpyHostToDevice, h->copy_stream));
	CU(cudaEventRecord(s.copied, h->copy_stream));
	CU(cudaStreamWaitEvent(h->stream, s.copied, 0));
	// offsets are relative to read_off[0] on the device: bias the base pointer instead of rewriting them
	const char* d_bases = (const char*)s.bases.p - first;
	rc = launch_map(h, d_bases, (const uint32_t*)s.off.p, (const uint32_t*)s.bc.p, n_pairs, j_index,
	    conreci_out ? (int32_t*)s.out.p : nullptr);
	if (rc)
		return rc;
	if (conreci_out)
		CU(cudaMemcpyAsync(conreci_out, s.out.p, n_pairs * 4ull, cudaMemcpyDeviceToHost, h->stream));
	CU(cudaEventRecord(s.done, h->stream));
	s.busy = true;
	s.copy_pending = true;
	s.wants_out = conreci_out != nullptr;
	return ARKS_OK;
}

int arks_map_pairs_device(arks_handle* h, const char* d_bases, const uint32_t* d_read_off, const uint32_t* d_barcode_id,
    uint32_t n_pairs, uint64_t n_bases, double j_index, int32_t* d_conreci_out)
{
	(void)n_bases;
	if (!h || !d_bases || !d_read_off || !d_barcode_id)
		return fail(h, ARKS_E_ARG, "arks_map_pairs_device: null argument");
	if (!h->finalized)
		return fail(h, ARKS_E_STATE, "arks_map_pairs before arks_index_finalize");
	if (n_pairs == 0)
		return ARKS_OK;
	CU(cudaSetDevice(h->device));
	int rc;
	if ((rc = imap_reserve(h, n_pairs)))
		return rc;
	return launch_map(h, d_bases, d_read_off, d_barcode_id, n_pairs, j_index, d_conreci_out);
}

int arks_map_get_stats(arks_handle* h, arks_map_stats* stats)
{
	if (!h || !stats)
		return ARKS_E_ARG;
	CU(cudaSetDevice(h->device));
	MapCounters c;
	CU(cudaMemcpyAsync(&c, h->d_mctr, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	stats->kmers_valid = c.kmers_valid;
	stats->kmers_invalid = c.kmers_invalid;
	stats->found = c.found;
	stats->recorded = c.recorded;
	stats->dups = c.dups;
	stats->reads_pass = c.reads_pass;
	stats->reads_fail = c.reads_fail;
	stats->pairs_stored = c.pairs_stored;
	stats->pairs_invalid = c.pairs_invalid;
	stats->pairs_nogood = c.pairs_nogood;
	if (c.overflow)
		return fail(h, ARKS_E_OVERFLOW, "a read hit more than 32 distinct contig ends; results for such reads are not exact");
	return ARKS_OK;
}

int arks_map_stats_reset(arks_handle* h)
{
	if (!h)
		return ARKS_E_ARG;
	CU(cudaSetDevice(h->device));
	CU(cudaMemsetAsync(h->d_mctr, 0, sizeof(MapCounters), h->stream));
	return ARKS_OK;
}

// ---- imap / pmap -------------------------------------------------------------------------

int arks_imap_size(arks_handle* h, uint64_t* n_rows)
{
	if (!h || !n_rows)
		return ARKS_E_ARG;
	CU(cudaSetDevice(h->device));
	unsigned long long n = 0;
	CU(cudaMemcpyAsync(&n, h->d_imap_count, 8, cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	*n_rows = n;
	return ARKS_OK;
}

int arks_imap_clear(arks_handle* h)
{
	if (!h)
		return ARKS_E_ARG;
	CU(cudaSetDevice(h->device));
	init_slots_kernel<<<grid_for(h, h->imap_cap * 2, 256, 8), 256, 0, h->stream>>>(h->imap, h->imap_cap * 2, 2);
	h->launches++;
	CU(cudaGetLastError());
	CU(cudaMemsetAsync(h->d_imap_count, 0, 8, h->stream));
	h->imap_upper = 0;
	h->pmap_ready = false;
	return ARKS_OK;
}

int arks_imap_export(arks_handle* h, uint32_t* barcode, uint32_t* contig, uint32_t* head, uint32_t* tail, uint64_t cap, uint64_t* n_rows)
{
	if (!h || !barcode || !contig || !head || !tail || !n_rows)
		return ARKS_E_ARG;
	int rc = arks_imap_size(h, n_rows);
	if (rc)
		return rc;
	const uint64_t n = *n_rows;
	if (cap < n)
		return fail(h, ARKS_E_ARG, "arks_imap_export: cap too small");
	if (n == 0)
		return ARKS_OK;
	uint32_t* d = nullptr;
	CU(cudaMalloc(&d, n * 16));
	CU(cudaMemsetAsync(h->d_scratch, 0, 8, h->stream));
	imap_export_kernel<<<grid_for(h, h->imap_cap, 256, 8), 256, 0, h->stream>>>(h->imap, h->imap_cap, d, d + n, d + 2 * n, d + 3 * n,
	    h->d_scratch, n);
	h->launches++;
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(barcode, d, n * 4, cudaMemcpyDeviceToHost, h->stream));
	CU(cudaMemcpyAsync(contig, d + n, n * 4, cudaMemcpyDeviceToHost, h->stream));
	CU(cudaMemcpyAsync(head, d + 2 * n, n * 4, cudaMemcpyDeviceToHost, h->stream));
	CU(cudaMemcpyAsync(tail, d + 3 * n, n * 4, cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	cudaFree(d);
	return ARKS_OK;
}

int arks_imap_add(arks_handle* h, const uint32_t* barcode, const uint32_t* contig, const uint32_t* head, const uint32_t* tail, uint64_t n_rows)
{
	if (!h || !barcode || !contig || !head || !tail)
		return ARKS_E_ARG;
	if (n_rows == 0)
		return ARKS_OK;
	CU(cudaSetDevice(h->device));
	int rc;
	if ((rc = imap_reserve(h, n_rows)))
		return rc;
	uint32_t* d = nullptr;
	CU(cudaMalloc(&d, n_rows * 16));
	CU(cudaMemcpyAsync(d, barcode, n_rows * 4, cudaMemcpyHostToDevice, h->stream));
	CU(cudaMemcpyAsync(d + n_rows, contig, n_rows * 4, cudaMemcpyHostToDevice, h->stream));
	CU(cudaMemcpyAsync(d + 2 * n_rows, head, n_rows * 4, cudaMemcpyHostToDevice, h->stream));
	CU(cudaMemcpyAsync(d + 3 * n_rows, tail, n_rows * 4, cudaMemcpyHostToDevice, h->stream));
	imap_add_rows_kernel<<<grid_for(h, n_rows, 256, 8), 256, 0, h->stream>>>(d, d + n_rows, d + 2 * n_rows, d + 3 * n_rows, n_rows,
	    h->imap, h->imap_cap - 1, h->d_imap_count);
	h->launches++;
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(h->stream));
	cudaFree(d);
	return ARKS_OK;
}

int arks_head_tail_table(int min_reads, float error_percent, uint32_t n, uint32_t* min_max)
{
	if (!min_max)
		return ARKS_E_ARG;
	return build_ht_table(min_reads, error_percent, n, min_max);
}

int arks_pair_links(arks_handle* h, const int32_t* mult, uint32_t n_barcodes, int min_mult, int max_mult, int min_reads,
    float error_percent, const uint32_t* lexrank, uint32_t n_contigs)
{
	if (!h || !mult || !lexrank)
		return fail(h, ARKS_E_ARG, "arks_pair_links: null argument");
	CU(cudaSetDevice(h->device));
	h->pmap_ready = false;
	h->pm_n = 0;
	cudaStream_t st = h->stream;
	for (uint32_t c = 0; c < n_contigs; ++c)
		if (lexrank[c] >= n_contigs)
			return fail(h, ARKS_E_ARG, "arks_pair_links: lexrank must be < n_contigs");
	// 1. largest head+tail -> decision table
	uint32_t* d_maxsum = reinterpret_cast<uint32_t*>(h->d_scratch);
	CU(cudaMemsetAsync(h->d_scratch, 0, 64, st));
	imap_maxsum_kernel<<<grid_for(h, h->imap_cap, 256, 8), 256, 0, st>>>(h->imap, h->imap_cap, d_maxsum);
	h->launches++;
	uint32_t maxsum = 0;
	CU(cudaMemcpyAsync(&maxsum, d_maxsum, 4, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	// the decision table depends on (min_reads, error_percent) only: kept and extended between calls
	if (h->ht_min_reads != min_reads || h->ht_error != error_percent) {
		h->ht_table.clear();
		h->ht_min_reads = min_reads;
		h->ht_error = error_percent;
	}
	int rc = ARKS_OK;
	if (h->ht_table.size() < (size_t)maxsum + 1) {
		const size_t want = std::max<size_t>((size_t)maxsum + 1, 2 * h->ht_table.size());
		std::vector<uint32_t> t(want);
		rc = build_ht_table(min_reads, error_percent, (uint32_t)want, t.data());
		if (rc)
			return fail(h, rc, "head/tail predicate is not monotone in max for this (min_reads, error_percent)");
		h->ht_table.swap(t);
	}
	const std::vector<uint32_t> table(h->ht_table.begin(), h->ht_table.begin() + maxsum + 1);
	const bool timing = getenv("ARKS_TIMING") != nullptr;
	auto tnow = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double t_last = tnow();
	auto lap = [&](const char* what) {
		if (timing) {
			cudaStreamSynchronize(st);
			const double t = tnow();
			fprintf(stderr, "arks_pair_links: %-34s %.3f ms\n", what, 1e3 * (t - t_last));
			t_last = t;
		}
	};
	lap("largest sum + decision table");
	const uint32_t nb = std::max<uint32_t>(n_barcodes, 1);
	const uint32_t nc = std::max<uint32_t>(n_contigs, 1);
	const uint32_t n_scan_blocks = (nb + kScanBlock - 1) / kScanBlock;
	if ((rc = ensure(h, h->lk_table, table.size() * 4)) || (rc = ensure(h, h->lk_mult, nb * 4ull)) || (rc = ensure(h, h->lk_cnt, nb * 4ull)) ||
	    (rc = ensure(h, h->lk_fill, nb * 4ull)) || (rc = ensure(h, h->lk_offs, (nb + 1) * 4ull)) ||
	    (rc = ensure(h, h->lk_sums, n_scan_blocks * 4ull)) || (rc = ensure(h, h->lk_rank, nc * 4ull)) || (rc = ensure(h, h->lk_inv, nc * 4ull)))
		return rc;
	uint32_t* d_table = (uint32_t*)h->lk_table.p;
	int32_t* d_mult = (int32_t*)h->lk_mult.p;
	uint32_t *d_cnt = (uint32_t*)h->lk_cnt.p, *d_fill = (uint32_t*)h->lk_fill.p, *d_offs = (uint32_t*)h->lk_offs.p,
	         *d_sums = (uint32_t*)h->lk_sums.p, *d_rank = (uint32_t*)h->lk_rank.p, *d_inv = (uint32_t*)h->lk_inv.p;
	CU(cudaMemcpyAsync(d_table, table.data(), table.size() * 4, cudaMemcpyHostToDevice, st));
	CU(cudaMemsetAsync(d_mult, 0, nb * 4ull, st));
	if (n_barcodes)
		CU(cudaMemcpyAsync(d_mult, mult, n_barcodes * 4ull, cudaMemcpyHostToDevice, st));
	CU(cudaMemsetAsync(d_inv, 0xFF, nc * 4ull, st));
	if (n_contigs) {
		CU(cudaMemcpyAsync(d_rank, lexrank, n_contigs * 4ull, cudaMemcpyHostToDevice, st));
		rank_inverse_kernel<<<grid_for(h, n_contigs, 256, 8), 256, 0, st>>>(d_rank, n_contigs, d_inv);
		h->launches++;
	}
	h->pm_n_contigs = n_contigs;
	CU(cudaMemsetAsync(d_cnt, 0, nb * 4ull, st));
	CU(cudaMemsetAsync(d_fill, 0, nb * 4ull, st));
	LinkParams L{h->imap, h->imap_cap, d_mult, n_barcodes, min_mult, max_mult, d_table, (uint32_t)table.size()};
	const int g_imap = grid_for(h, h->imap_cap, 256, 8);
	// 2. count passing rows per barcode, 3. scan, 4. scatter (rows grouped by barcode)
	imap_count_kernel<<<g_imap, 256, 0, st>>>(L, d_cnt);
	scan_block_sums_kernel<<<n_scan_blocks, kScanBlock, 0, st>>>(d_cnt, nb, d_sums);
	scan_sums_kernel<<<1, kScanBlock, 0, st>>>(d_sums, n_scan_blocks);
	scan_apply_kernel<<<n_scan_blocks, kScanBlock, 0, st>>>(d_cnt, nb, d_sums, d_offs);
	unsigned long long* d_events = h->d_scratch + 1;
	pair_count_kernel<<<grid_for(h, nb, 256, 8), 256, 0, st>>>(d_cnt, nb, d_events);
	h->launches += 5;
	CU(cudaGetLastError());
	uint32_t n_rows = 0;
	unsigned long long events = 0;
	CU(cudaMemcpyAsync(&n_rows, d_offs + nb, 4, cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(&events, d_events, 8, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	if ((rc = ensure(h, h->lk_rows, std::max<uint32_t>(n_rows, 1) * 4ull)) || (rc = ensure(h, h->lk_rowbc, std::max<uint32_t>(n_rows, 1) * 4ull)))
		return rc;
	uint32_t *d_rows = (uint32_t*)h->lk_rows.p, *d_rowbc = (uint32_t*)h->lk_rowbc.p;
	lap("count rows per barcode + scan");
	imap_scatter_kernel<<<g_imap, 256, 0, st>>>(L, d_offs, d_fill, d_rows, d_rowbc);
	h->launches++;
	CU(cudaGetLastError());
	lap("rows grouped by barcode");
	// 5. the pmap hash.  The number of pair events bounds the number of distinct pairs from above, usually by
	// far (neighbouring contigs share many barcodes), so the table starts at 2 x min(events, all pairs, 2^25) slots
	// and is doubled -- and the pass repeated -- if it fills up beyond 70 %.
	unsigned long long bound = events;
	if (n_contigs)
		bound = std::min(bound, (unsigned long long)n_contigs * (n_contigs - 1) / 2);
	uint64_t cap = 1024;
	while (cap < 2 * std::min<unsigned long long>(bound, 1ull << 25))
		cap <<= 1;
	if (const char* e = getenv("ARKS_PMAP_INITIAL_SLOTS")) { // tests: force the growth path
		cap = 1024;
		while (cap < (uint64_t)atoll(e))
			cap <<= 1;
	}
	unsigned long long n_pairs = 0;
	uint32_t* d_overflow = reinterpret_cast<uint32_t*>(h->d_scratch + 2);
	while (true) {
		if ((rc = ensure(h, h->pmap_buf, cap * 32)))
			return rc;
		h->pmap_cap = cap;
		unsigned long long* pmap = (unsigned long long*)h->pmap_buf.p;
		init_slots_kernel<<<grid_for(h, cap * 4, 256, 8), 256, 0, st>>>(pmap, cap * 4, 4);
		h->launches++;
		CU(cudaMemsetAsync(h->d_pmap_count, 0, 8, st));
		CU(cudaMemsetAsync(d_overflow, 0, 4, st));
		const uint64_t limit = cap >= 2 * bound ? cap : cap / 10 * 7;
		if (events && n_rows) {
			pair_kernel<<<grid_for(h, n_rows, 128, 16), 128, 0, st>>>(d_offs, d_rows, d_rowbc, n_rows, d_rank, pmap, cap - 1, h->d_pmap_count,
			    limit, d_overflow);
			h->launches++;
		}
		CU(cudaGetLastError());
		uint32_t overflow = 0;
		CU(cudaMemcpyAsync(&n_pairs, h->d_pmap_count, 8, cudaMemcpyDeviceToHost, st));
		CU(cudaMemcpyAsync(&overflow, d_overflow, 4, cudaMemcpyDeviceToHost, st));
		CU(cudaStreamSynchronize(st));
		if (!overflow)
			break;
		cap <<= 1;
	}
	if (timing)
		fprintf(stderr, "arks_pair_links: %u rows, %llu pair events, %llu distinct pairs, %llu pmap slots\n", n_rows, events, n_pairs,
		    (unsigned long long)cap);
	lap("pair kernel into the pmap hash");
	// 6. order by (rank a, rank b) = std::map<pair<string,string>> iteration order, on the device
	h->pm_n = n_pairs;
	const uint64_t n1 = std::max<uint64_t>(n_pairs, 1);
	if ((rc = ensure(h, h->st_key[0], n1 * 8)) || (rc = ensure(h, h->st_val[0], n1 * 4)) || (rc = ensure(h, h->pm_keys, n1 * 8)) ||
	    (rc = ensure(h, h->pm_counts, n1 * 16)))
		return rc;
	if (n_pairs) {
		CU(cudaMemsetAsync(h->d_scratch, 0, 8, st));
		pmap_collect_kernel<<<grid_for(h, cap, 256, 8), 256, 0, st>>>((const unsigned long long*)h->pmap_buf.p, cap,
		    (unsigned long long*)h->st_key[0].p, (uint32_t*)h->st_val[0].p, h->d_scratch, n_pairs);
		h->launches++;
		int buf = 0;
		if ((rc = device_sort_pairs(h, n_pairs, bits_for(n_contigs), &buf)))
			return rc;
		CU(cudaMemcpyAsync(h->pm_keys.p, h->st_key[buf].p, n_pairs * 8, cudaMemcpyDeviceToDevice, st));
		pmap_gather_counts_kernel<<<grid_for(h, n_pairs, 256, 8), 256, 0, st>>>((const unsigned long long*)h->pmap_buf.p,
		    (const uint32_t*)h->st_val[buf].p, n_pairs, (uint32_t*)h->pm_counts.p);
		h->launches++;
		CU(cudaGetLastError());
	}
	if ((rc = pmap_fill_names(h)))
		return rc;
	CU(cudaStreamSynchronize(st));
	lap("collect + radix sort + gather");
	h->pmap_ready = true;
	return ARKS_OK;
}

int arks_pmap_size(arks_handle* h, uint64_t* n_rows)
{
	if (!h || !n_rows)
		return ARKS_E_ARG;
	if (!h->pmap_ready)
		return fail(h, ARKS_E_STATE, "arks_pmap_size before arks_pair_links");
	*n_rows = h->pm_n;
	return ARKS_OK;
}

int arks_pmap_export(arks_handle* h, uint32_t* a, uint32_t* b, uint32_t* counts4, uint64_t cap, uint64_t* n_rows)
{
	if (!h || !a || !b || !counts4 || !n_rows)
		return ARKS_E_ARG;
	if (!h->pmap_ready)
		return fail(h, ARKS_E_STATE, "arks_pmap_export before arks_pair_links");
	const uint64_t n = h->pm_n;
	*n_rows = n;
	if (cap < n)
		return fail(h, ARKS_E_ARG, "arks_pmap_export: cap too small");
	if (n) {
		CU(cudaSetDevice(h->device));
		// the rows are already in order on the device: three copies (at PCIe rate when the caller's arrays
		// come from arks_host_alloc)
		CU(cudaMemcpyAsync(a, h->pm_a.p, n * 4, cudaMemcpyDeviceToHost, h->stream));
		CU(cudaMemcpyAsync(b, h->pm_b.p, n * 4, cudaMemcpyDeviceToHost, h->stream));
		CU(cudaMemcpyAsync(counts4, h->pm_counts.p, n * 16, cudaMemcpyDeviceToHost, h->stream));
		CU(cudaStreamSynchronize(h->stream));
	}
	return ARKS_OK;
}

int arks_pmap_digest(arks_handle* h, uint64_t digest[2])
{
	if (!h || !digest)
		return ARKS_E_ARG;
	if (!h->pmap_ready)
		return fail(h, ARKS_E_STATE, "arks_pmap_digest before arks_pair_links");
	CU(cudaSetDevice(h->device));
	CU(cudaMemsetAsync(h->d_scratch + 4, 0, 16, h->stream));
	if (h->pm_n) {
		pmap_digest_kernel<<<grid_for(h, h->pm_n, 256, 8), 256, 0, h->stream>>>((const unsigned long long*)h->pm_keys.p,
		    (const uint32_t*)h->pm_counts.p, h->pm_n, h->d_scratch + 4);
		h->launches++;
		CU(cudaGetLastError());
	}
	unsigned long long d[2];
	CU(cudaMemcpyAsync(d, h->d_scratch + 4, 16, cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	digest[0] = d[0];
	digest[1] = d[1];
	return ARKS_OK;
}

// ---- multi-GPU: one exchange step for the pair-link map ---------------------------------------------

int arks_comm_unique_id(uint8_t id[ARKS_COMM_ID_BYTES])
{
	arks_handle* h = nullptr;
	if (!id)
		return ARKS_E_ARG;
	const NcclApi& N = nccl_api();
	if (!N.ok)
		return fail(h, ARKS_E_NCCL, N.error);
	static_assert(sizeof(ncclUniqueId) == ARKS_COMM_ID_BYTES, "ncclUniqueId size");
	ncclUniqueId uid;
	ncclResult_t r = N.GetUniqueId(&uid);
	if (r != ncclSuccess)
		return fail(h, ARKS_E_NCCL, std::string("ncclGetUniqueId: ") + N.GetErrorString(r));
	memcpy(id, &uid, sizeof(uid));
	return ARKS_OK;
}

int arks_comm_init_rank(arks_handle* h, const uint8_t id[ARKS_COMM_ID_BYTES], int rank, int n_ranks)
{
	if (!h || !id || rank < 0 || rank >= n_ranks)
		return fail(h, ARKS_E_ARG, "arks_comm_init_rank: bad argument");
	const NcclApi& N = nccl_api();
	if (!N.ok)
		return fail(h, ARKS_E_NCCL, N.error);
	CU(cudaSetDevice(h->device));
	comm_release(h);
	ncclUniqueId uid;
	memcpy(&uid, id, sizeof(uid));
	ncclComm_t comm = nullptr;
	ncclResult_t r = N.CommInitRank(&comm, n_ranks, uid, rank);
	if (r != ncclSuccess)
		return fail(h, ARKS_E_NCCL, std::string("ncclCommInitRank: ") + N.GetErrorString(r));
	h->nccl_comm = comm;
	h->comm_rank = rank;
	h->comm_size = n_ranks;
	return ARKS_OK;
}

int arks_comm_init_local(arks_handle** hs, int n)
{
	if (!hs || n < 1)
		return ARKS_E_ARG;
	arks_handle* h = hs[0];
	for (int i = 0; i < n; ++i)
		if (!hs[i])
			return fail(h, ARKS_E_ARG, "arks_comm_init_local: null handle");
	std::vector<int> devs(n);
	bool distinct = true;
	for (int i = 0; i < n; ++i) {
		devs[i] = hs[i]->device;
		for (int j = 0; j < i; ++j)
			distinct &= devs[j] != devs[i];
	}
	for (int i = 0; i < n; ++i) {
		comm_release(hs[i]);
		hs[i]->comm_rank = i;
		hs[i]->comm_size = n;
	}
	if (n == 1 || !distinct)
		return ARKS_OK; // shards that share a device exchange with device-to-device copies (arks_merge_pmap)
	const NcclApi& N = nccl_api();
	if (!N.ok)
		return fail(h, ARKS_E_NCCL, N.error);
	std::vector<ncclComm_t> comms(n, nullptr);
	ncclResult_t r = N.CommInitAll(comms.data(), n, devs.data());
	if (r != ncclSuccess)
		return fail(h, ARKS_E_NCCL, std::string("ncclCommInitAll: ") + N.GetErrorString(r));
	for (int i = 0; i < n; ++i)
		hs[i]->nccl_comm = comms[i];
	return ARKS_OK;
}

// The pair-link maps of barcode-disjoint shards add up key by key.  Every handle ends up with the
// merged map: all-gather of the sorted keys (sizes first) -> sort + unique of the concatenation = the same sorted
// union on every rank -> this rank's counters scattered into a dense 4 x n_union vector -> ONE all-reduce (sum,
// uint32) -> the union and the reduced vector become the handle's map.  Integer sums: the result depends neither
// on the number of shards nor on arrival order.
int arks_merge_pmap(arks_handle** hs, int n_local)
{
	if (!hs || n_local < 1 || !hs[0])
		return ARKS_E_ARG;
	arks_handle* h = hs[0];
	const int world = h->comm_size;
	for (int i = 0; i < n_local; ++i) {
		if (!hs[i] || !hs[i]->pmap_ready)
			return fail(h, ARKS_E_STATE, "arks_merge_pmap before arks_pair_links");
		if (hs[i]->comm_size != world || hs[i]->pm_n_contigs != h->pm_n_contigs)
			return fail(h, ARKS_E_ARG, "arks_merge_pmap: handles of different communicators / contig sets");
	}
	if (world == 1)
		return ARKS_OK;
	const bool use_nccl = h->nccl_comm != nullptr;
	if (!use_nccl && n_local != world)
		return fail(h, ARKS_E_STATE, "arks_merge_pmap: no communicator (arks_comm_init_rank / arks_comm_init_local first)");
	const NcclApi& N = nccl_api();
#define NC(call)                                                                                         \
	do {                                                                                                 \
		ncclResult_t r_ = (call);                                                                        \
		if (r_ != ncclSuccess)                                                                           \
			return fail(h, ARKS_E_NCCL, std::string(#call) + ": " + N.GetErrorString(r_));               \
	} while (0)
	// ---- sizes of every rank's map
	std::vector<unsigned long long> sizes(world, 0);
	if (n_local == world) {
		for (int i = 0; i < n_local; ++i)
			sizes[hs[i]->comm_rank] = hs[i]->pm_n;
	} else {
		for (int i = 0; i < n_local; ++i) {
			arks_handle* g = hs[i];
			CU(cudaSetDevice(g->device));
			int rc = ensure(g, g->mg_head, (uint64_t)world * 8);
			if (rc)
				return rc;
			const unsigned long long mine = g->pm_n;
			CU(cudaMemcpyAsync(g->d_scratch + 3, &mine, 8, cudaMemcpyHostToDevice, g->stream));
			CU(cudaStreamSynchronize(g->stream)); // `mine` is a stack variable
		}
		NC(N.GroupStart());
		for (int i = 0; i < n_local; ++i)
			NC(N.AllGather(hs[i]->d_scratch + 3, hs[i]->mg_head.p, 1, ncclUint64, (ncclComm_t)hs[i]->nccl_comm, hs[i]->stream));
		NC(N.GroupEnd());
		CU(cudaSetDevice(h->device));
		CU(cudaMemcpyAsync(sizes.data(), h->mg_head.p, (size_t)world * 8, cudaMemcpyDeviceToHost, h->stream));
		CU(cudaStreamSynchronize(h->stream));
	}
	std::vector<uint64_t> off(world + 1, 0);
	for (int r = 0; r < world; ++r)
		off[r + 1] = off[r] + sizes[r];
	const uint64_t total = off[world];
	if (total == 0)
		return ARKS_OK;
	// ---- all-gather (v) of the sorted keys into the sort buffer of every handle
	for (int i = 0; i < n_local; ++i) {
		arks_handle* g = hs[i];
		CU(cudaSetDevice(g->device));
		int rc;
		if ((rc = ensure(g, g->st_key[0], total * 8)) || (rc = ensure(g, g->st_val[0], total * 4)))
			return rc;
	}
	if (use_nccl) {
		NC(N.GroupStart());
		for (int i = 0; i < n_local; ++i) {
			arks_handle* g = hs[i];
			for (int r = 0; r < world; ++r) {
				if (!sizes[r])
					continue;
				unsigned long long* dst = (unsigned long long*)g->st_key[0].p + off[r];
				NC(N.Broadcast(r == g->comm_rank ? g->pm_keys.p : (const void*)dst, dst, sizes[r], ncclUint64, r, (ncclComm_t)g->nccl_comm,
				    g->stream));
			}
		}
		NC(N.GroupEnd());
	} else {
		// all shards live in this process on one device: plain device-to-device copies
		for (int i = 0; i < n_local; ++i)
			CU(cudaStreamSynchronize(hs[i]->stream));
		for (int i = 0; i < n_local; ++i) {
			arks_handle* g = hs[i];
			CU(cudaSetDevice(g->device));
			for (int j = 0; j < n_local; ++j) {
				const int r = hs[j]->comm_rank;
				if (sizes[r])
					CU(cudaMemcpyAsync((unsigned long long*)g->st_key[0].p + off[r], hs[j]->pm_keys.p, sizes[r] * 8, cudaMemcpyDeviceToDevice,
					    g->stream));
			}
		}
	}
	// ---- the sorted union, identically on every handle
	std::vector<int> bufs(n_local, 0);
	for (int i = 0; i < n_local; ++i) {
		arks_handle* g = hs[i];
		CU(cudaSetDevice(g->device));
		int rc;
		if ((rc = device_sort_pairs(g, total, bits_for(g->pm_n_contigs), &bufs[i])))
			return rc;
		const uint32_t n_scan_blocks = (uint32_t)((total + kScanBlock - 1) / kScanBlock);
		if ((rc = ensure(g, g->mg_head, (total + 1) * 4)) || (rc = ensure(g, g->st_sums, (uint64_t)n_scan_blocks * 4)) ||
		    (rc = ensure(g, g->mg_all, total * 8)))
			return rc;
		const unsigned long long* sorted = (const unsigned long long*)g->st_key[bufs[i]].p;
		uint32_t* head = (uint32_t*)g->mg_head.p;
		run_heads_kernel<<<grid_for(g, total, 256, 8), 256, 0, g->stream>>>(sorted, total, head);
		scan_block_sums_kernel<<<n_scan_blocks, kScanBlock, 0, g->stream>>>(head, total, (uint32_t*)g->st_sums.p);
		scan_sums_kernel<<<1, kScanBlock, 0, g->stream>>>((uint32_t*)g->st_sums.p, n_scan_blocks);
		scan_apply_kernel<<<n_scan_blocks, kScanBlock, 0, g->stream>>>(head, total, (const uint32_t*)g->st_sums.p, head);
		compact_unique_kernel<<<grid_for(g, total, 256, 8), 256, 0, g->stream>>>(sorted, total, head, (unsigned long long*)g->mg_all.p);
		g->launches += 5;
		CU(cudaGetLastError());
	}
	uint32_t n_uni = 0;
	CU(cudaSetDevice(h->device));
	CU(cudaMemcpyAsync(&n_uni, (uint32_t*)h->mg_head.p + total, 4, cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	// ---- dense counter vector, one all-reduce
	for (int i = 0; i < n_local; ++i) {
		arks_handle* g = hs[i];
		CU(cudaSetDevice(g->device));
		int rc;
		if ((rc = ensure(g, g->mg_dense, (uint64_t)n_uni * 16)))
			return rc;
		CU(cudaMemsetAsync(g->mg_dense.p, 0, (uint64_t)n_uni * 16, g->stream));
		if (g->pm_n) {
			scatter_counts_kernel<<<grid_for(g, g->pm_n, 256, 8), 256, 0, g->stream>>>((const unsigned long long*)g->pm_keys.p,
			    (const uint32_t*)g->pm_counts.p, g->pm_n, (const unsigned long long*)g->mg_all.p, n_uni, (uint32_t*)g->mg_dense.p);
			g->launches++;
			CU(cudaGetLastError());
		}
	}
	if (use_nccl) {
		NC(N.GroupStart());
		for (int i = 0; i < n_local; ++i)
			NC(N.AllReduce(hs[i]->mg_dense.p, hs[i]->mg_dense.p, (size_t)n_uni * 4, ncclUint32, ncclSum, (ncclComm_t)hs[i]->nccl_comm,
			    hs[i]->stream));
		NC(N.GroupEnd());
	} else {
		for (int i = 0; i < n_local; ++i)
			CU(cudaStreamSynchronize(hs[i]->stream));
		CU(cudaSetDevice(h->device));
		for (int j = 1; j < n_local; ++j) {
			add_u32_kernel<<<grid_for(h, (uint64_t)n_uni * 4, 256, 8), 256, 0, h->stream>>>((uint32_t*)h->mg_dense.p,
			    (const uint32_t*)hs[j]->mg_dense.p, (uint64_t)n_uni * 4);
			h->launches++;
		}
		CU(cudaGetLastError());
		CU(cudaStreamSynchronize(h->stream));
		for (int j = 1; j < n_local; ++j)
			CU(cudaMemcpyAsync(hs[j]->mg_dense.p, h->mg_dense.p, (uint64_t)n_uni * 16, cudaMemcpyDeviceToDevice, hs[j]->stream));
	}
	// ---- the union and the reduced vector become every handle's map
	for (int i = 0; i < n_local; ++i) {
		arks_handle* g = hs[i];
		CU(cudaSetDevice(g->device));
		std::swap(g->pm_keys, g->mg_all);
		std::swap(g->pm_counts, g->mg_dense);
		g->pm_n = n_uni;
		int rc = pmap_fill_names(g);
		if (rc)
			return rc;
	}
	for (int i = 0; i < n_local; ++i) {
		CU(cudaSetDevice(hs[i]->device));
		CU(cudaStreamSynchronize(hs[i]->stream));
	}
#undef NC
	return ARKS_OK;
}

} // extern "C"
