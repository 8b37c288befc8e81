/* include/arks_b200.h -- C ABI of libarks_b200.so
 *
 * B200-native (sm_100a) implementation of the ARKS hot path of bcgsc/arcs 1.2.8:
 * contig-end k-merisation into an exact GPU hash table, per-read-pair k-mer lookup
 * + contig-end vote, per-barcode tallies and pairwise link counters.
 *
 * The reference has no FFI of its own: the seams replaced are three internal C++
 * calls of runArcs (Arcs/Arcs.cpp:1871-1909).  Each entry point cites the one it
 * replaces.  Plain pointers and sizes only; one handle per GPU; a handle is not
 * re-entrant (drive different handles from different threads).
 *
 * All functions return ARKS_OK (0) or a negative error; arks_last_error() gives text.
 * There is NO CPU fallback: without a CUDA device arks_create() fails.
 */
#ifndef ARKS_B200_H
#define ARKS_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ARKS_OK 0
#define ARKS_E_CUDA -1      /* a CUDA runtime call failed */
#define ARKS_E_ARG -2       /* bad argument (k out of range, null pointer, ...) */
#define ARKS_E_STATE -3     /* call order violated (e.g. map before index_finalize) */
#define ARKS_E_CAPACITY -4  /* index table full: more distinct k-mers than max_kmers */
#define ARKS_E_OVERFLOW -5  /* a read hit more than ARKS_MAX_TRACK distinct contig ends */
#define ARKS_E_NOMEM -6
#define ARKS_E_NONMONOTONE -7 /* head/tail predicate not monotone in max for some sum */
#define ARKS_E_NCCL -8      /* NCCL could not be loaded or one of its calls failed */

#define ARKS_MIN_K 4   /* ReadsProcessor requires k > 3 (Common/ReadsProcessor.cpp:25) */
#define ARKS_MAX_K 64  /* 128-bit device keys */

typedef struct arks_handle arks_handle;

/* Counters of getContigKmers' verbose block (Arcs/Arcs.cpp:179-180,1107-1128). */
typedef struct
{
	uint64_t kmers_valid; /* "Total number of Kmers": valid windows visited by the mapKmers walk */
	uint64_t kmers_null;  /* s_numbadkmers: NULL windows visited (each skips k positions) */
	uint64_t recorded;    /* s_numkmersmapped: distinct keys */
	uint64_t collisions;  /* s_numkmercollisions = kmers_valid - recorded */
	uint64_t removed;     /* s_numkmersremdup */
	uint64_t unique;      /* s_uniquedraftkmers: keys seen in exactly one contig end */
} arks_index_stats;

/* Counters of chromiumRead's verbose block (Arcs/Arcs.cpp:182-185,1321-1340). */
typedef struct
{
	uint64_t kmers_valid;   /* s_totalnumckmers */
	uint64_t kmers_invalid; /* s_numbadckmers */
	uint64_t found;         /* s_numckmersfound */
	uint64_t recorded;      /* s_numckmersrec */
	uint64_t dups;          /* s_ckmersasdups */
	uint64_t reads_pass;    /* s_numreadspassingjaccard */
	uint64_t reads_fail;    /* s_numreadsfailjaccard */
	uint64_t pairs_stored;  /* stored_readpairs */
	uint64_t pairs_invalid; /* skipped_invalidreadpair */
	uint64_t pairs_nogood;  /* skipped_nogoodcontig */
} arks_map_stats;

/* ---- lifetime ------------------------------------------------------------------ */

/* device: CUDA ordinal.  k: k-mer size (ARKS_MIN_K..ARKS_MAX_K).  max_kmers: upper bound
 * on the number of contig-end k-mer windows that will be added (sizes the table; the
 * table holds distinct keys only, so the sum of end lengths is always enough).
 * Replaces: ContigKMap kmap + ReadsProcessor proc(k) (Arcs/Arcs.cpp:1850-1851,1041-1042). */
int arks_create(int device, int k, uint64_t max_kmers, arks_handle** out);
void arks_destroy(arks_handle* h);
const char* arks_last_error(const arks_handle* h); /* h may be NULL: last create error */
/* Run all kernels of this handle on an existing CUDA stream (cudaStream_t passed as
 * void*; NULL = the handle's own stream).  Lets a caller time with its own events. */
int arks_set_stream(arks_handle* h, void* cuda_stream);
int arks_sync(arks_handle* h);
/* Pinned host memory for the batch buffers passed to arks_index_add / arks_map_pairs. */
int arks_host_alloc(void** p, size_t bytes);
/* Creates the CUDA context of `device` (the half second to a second every CUDA process pays once).
 * Optional: a host that has other start-up work (parsing the draft) can call it from a helper
 * thread first so that arks_create does not wait for it. */
int arks_device_init(int device);
int arks_host_free(void* p);
/* Multi-GPU hosts: pins the CALLING thread to the cores of the NUMA node the device hangs off
 * (sysfs topology; a no-op when it cannot be read), so that the pinned buffers the thread
 * allocates afterwards (first touch) and its copies stay off the inter-socket link.
 * arks_device_numa_node reports that node (-1: unknown). */
int arks_bind_thread(int device);
int arks_device_numa_node(int device, int* node);

/* ---- kernel 1: contig-end k-mer index -------------------------------------------- */

/* Adds n_ends contig ends.  End e is bases[end_off[e] .. end_off[e+1]) (ASCII, any
 * case, IUPAC allowed) and gets contig-end record conreci[e] >= 1 (head = 2i-1,
 * tail = 2i for the i-th kept contig; 0 is reserved for "seen in several ends").
 * Performs mapKmers' walk (NULL window => skip k positions) and its insert rule.
 * Host buffers.  May be called repeatedly.
 * Replaces: mapKmers (Arcs/Arcs.cpp:869-929) as called by getContigKmers (:1084-1091). */
int arks_index_add(arks_handle* h, const char* bases, const uint64_t* end_off, const uint32_t* conreci,
    uint32_t n_ends);
/* Same with device-resident inputs (no copies). */
int arks_index_add_device(arks_handle* h, const char* d_bases, const uint64_t* d_end_off,
    const uint32_t* d_conreci, const uint64_t* h_end_off, uint32_t n_ends);
/* Freezes the table (collapses build bookkeeping into the final value per key) and
 * returns the counters.  stats may be NULL. */
int arks_index_finalize(arks_handle* h, arks_index_stats* stats);
/* Number of distinct keys / copy of the table as (key bytes, value) rows in arbitrary
 * order.  Key bytes are exactly ReadsProcessor::getStr's ceil(k/4) bytes.  For tests. */
int arks_index_size(arks_handle* h, uint64_t* n_keys);
int arks_index_dump(arks_handle* h, uint8_t* keys, int32_t* values, uint64_t cap, uint64_t* n_keys);

/* ---- kernel 2: read-pair lookup + vote -------------------------------------------- */

/* Optional: remap[c] for c in [0, n) is the contig-end record under which hits on
 * conreci c are tallied (used to merge contigs that share a FASTA name, because the
 * reference's imap is keyed by name: Arcs/Arcs.cpp:1281-1284).  Default identity. */
int arks_set_conreci_remap(arks_handle* h, const uint32_t* remap, uint32_t n);

/* Maps n_pairs read pairs that already passed the name / barcode checks of
 * chromiumRead.  Read r is bases[read_off[r] .. read_off[r+1]) (ASCII); reads 2i and
 * 2i+1 are mates with barcode id barcode_id[i].  For every pair: checkReadSequence on
 * both reads, bestContig on both, and if both agree on a non-null contig end the
 * (barcode, contig end) tally is incremented on the device.
 * conreci_out (optional, host, n_pairs): the stored contig end per pair or 0.
 * Host buffers; returns once the inputs have been consumed (copied to the device);
 * the kernel itself runs asynchronously -- arks_sync() / arks_map_get_stats() wait.
 * Replaces: the body of chromiumRead's parallel loop, Arcs/Arcs.cpp:1266-1292, i.e.
 * checkReadSequence (:366-389) + bestContig (:939-1014) + imap[barcode][end]++. */
int arks_map_pairs(arks_handle* h, const char* bases, const uint32_t* read_off, const uint32_t* barcode_id,
    uint32_t n_pairs, double j_index, int32_t* conreci_out);
/* The same call in two halves, for a host that feeds several GPUs from one thread:
 * arks_map_pairs_begin enqueues the copies and the kernels and returns at once;
 * arks_map_pairs_end waits until the inputs of every begin since the last end have been
 * consumed (and, if conreci_out was given, until the results are there).  The input
 * buffers must not be touched in between.  arks_map_pairs == begin + end. */
int arks_map_pairs_begin(arks_handle* h, const char* bases, const uint32_t* read_off, const uint32_t* barcode_id,
    uint32_t n_pairs, double j_index, int32_t* conreci_out);
int arks_map_pairs_end(arks_handle* h);
/* Same with device-resident inputs and (optional) device output; fully asynchronous. */
int arks_map_pairs_device(arks_handle* h, const char* d_bases, const uint32_t* d_read_off,
    const uint32_t* d_barcode_id, uint32_t n_pairs, uint64_t n_bases, double j_index, int32_t* d_conreci_out);
int arks_map_get_stats(arks_handle* h, arks_map_stats* stats); /* synchronises */
int arks_map_stats_reset(arks_handle* h);

/* ---- per-barcode tallies (imap) and pair links (pmap) ------------------------------- */

/* imap as rows (barcode id, contig index, head count, tail count), one row per
 * (barcode, contig) with head+tail > 0, arbitrary order.  contig index = (conreci-1)/2.
 * Replaces: ARCS::IndexMap (Arcs/Arcs.h:106-113) after the zero-fill of
 * Arcs/Arcs.cpp:1309-1319 (a missing end is a 0 in the row). */
int arks_imap_size(arks_handle* h, uint64_t* n_rows);
int arks_imap_export(arks_handle* h, uint32_t* barcode, uint32_t* contig, uint32_t* head, uint32_t* tail,
    uint64_t cap, uint64_t* n_rows);
/* Forgets every tally (an empty IndexMap); the index and the map counters stay. */
int arks_imap_clear(arks_handle* h);
/* Adds rows to the device imap (used to merge tallies produced elsewhere, e.g. by
 * another GPU that saw part of a barcode, or by ARCS alignment mode). */
int arks_imap_add(arks_handle* h, const uint32_t* barcode, const uint32_t* contig, const uint32_t* head,
    const uint32_t* tail, uint64_t n_rows);

/* pairContigs on the device imap.  mult[b]: multiplicity of barcode id b (n_barcodes
 * entries); lexrank[c]: rank of contig c's name under std::string '<' (n_contigs
 * entries).  Builds the pair-link map on the device.
 * Replaces: pairContigs + headOrTail + normalEstimation (Arcs/Arcs.cpp:833-861,1378-1435). */
int arks_pair_links(arks_handle* h, const int32_t* mult, uint32_t n_barcodes, int min_mult, int max_mult,
    int min_reads, float error_percent, const uint32_t* lexrank, uint32_t n_contigs);
/* pmap rows (contig a, contig b, counts[4] = HH,HT,TH,TT) with lexrank[a] < lexrank[b],
 * sorted by (lexrank[a], lexrank[b]) = the iteration order of ARCS::PairMap (Arcs.h:115).
 * The rows are ordered on the device (radix sort); the export is three device->host copies,
 * at PCIe rate when a / b / counts4 come from arks_host_alloc. */
int arks_pmap_size(arks_handle* h, uint64_t* n_rows);
int arks_pmap_export(arks_handle* h, uint32_t* a, uint32_t* b, uint32_t* counts4, uint64_t cap,
    uint64_t* n_rows);
/* Order-independent 128-bit digest of the map's rows (computed on the device): equal maps
 * have equal digests whatever the number of GPUs they were accumulated on. */
int arks_pmap_digest(arks_handle* h, uint64_t digest[2]);

/* ---- multi-GPU: read pairs sharded by barcode, one exchange step ------------------------ */

/* Barcodes are the unit of independence of chromiumRead + pairContigs: with the read pairs
 * partitioned by barcode over G handles (every handle holding the same index), the
 * IndexMap shards are disjoint and the PairMap of the run is the key-wise SUM of the
 * shards' PairMaps.  arks_merge_pmap performs that sum over NCCL: all-gather of the
 * shards' sorted keys -> the same sorted union everywhere -> ONE ncclAllReduce (sum,
 * uint32) over the dense 4 x n_union counter vector.  Afterwards every handle's
 * arks_pmap_size / arks_pmap_export / arks_pmap_digest describe the merged map.
 * Replaces: nothing upstream (the reference is single-node, Arcs/Arcs.cpp:1378-1435 runs
 * on one IndexMap); the result equals pairContigs on the union of the shards.
 *
 * One process per GPU: rank 0 calls arks_comm_unique_id, the id travels to the other
 * ranks by any means, every rank calls arks_comm_init_rank with its handle, then -- after
 * arks_pair_links -- arks_merge_pmap(&h, 1) collectively.
 * One process, G handles: arks_comm_init_local(handles, G), then
 * arks_merge_pmap(handles, G) from one thread.  (Handles that share a device -- tests --
 * exchange with device-to-device copies instead of NCCL.) */
#define ARKS_COMM_ID_BYTES 128
int arks_comm_unique_id(uint8_t id[ARKS_COMM_ID_BYTES]);
int arks_comm_init_rank(arks_handle* h, const uint8_t id[ARKS_COMM_ID_BYTES], int rank, int n_ranks);
int arks_comm_init_local(arks_handle** handles, int n);
int arks_merge_pmap(arks_handle** handles, int n_local);

/* The exact head/tail decision table used by arks_pair_links: for sum in [0, n):
 * min_max[sum] = smallest max(head,tail) for which headOrTail() is valid, or
 * UINT32_MAX if none (computed on the host with the reference's float/double
 * expression).  For tests. */
int arks_head_tail_table(int min_reads, float error_percent, uint32_t n, uint32_t* min_max);

/* Number of kernel launches issued by this handle so far (bench.py's gpu_launches). */
uint64_t arks_launch_count(const arks_handle* h);

#ifdef __cplusplus
}
#endif
#endif
