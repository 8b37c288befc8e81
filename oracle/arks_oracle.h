/* oracle/arks_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the ARKS hot path of bcgsc/arcs 1.2.8, function by
 * function, used as the parity checker for the CUDA path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the
 * product (arcs_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks it against
 *   - the reference's unmodified ReadsProcessor::prepSeq (oracle/_ref/libref_prepseq.so),
 *   - the golden outputs of Examples/arks_test-demo and Examples/arks-long_test-demo,
 *   - dumps (kmap / per-read conreci / imap / pmap) of the reference's own code
 *     (oracle/_ref/arcs_ref) committed under tests/golden/.
 */
#ifndef ARKS_ORACLE_H
#define ARKS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ARKS_ORACLE_MAX_KEY_BYTES 64 /* k <= 256 */

typedef struct arks_okmap arks_okmap;

/* Counters of Arcs.cpp:179-185 (64-bit here; the reference's are 32-bit). */
typedef struct
{
	uint64_t kmers_valid;   /* valid windows visited  ("Total number of Kmers") */
	uint64_t kmers_null;    /* s_numbadkmers: invalid windows visited, each skipping k */
	uint64_t recorded;      /* s_numkmersmapped: distinct keys */
	uint64_t collisions;    /* s_numkmercollisions */
	uint64_t removed;       /* s_numkmersremdup (order dependent) */
	int64_t unique;         /* s_uniquedraftkmers */
} arks_oracle_index_stats;

typedef struct
{
	uint64_t kmers_valid;    /* s_totalnumckmers */
	uint64_t kmers_invalid;  /* s_numbadckmers */
	uint64_t found;          /* s_numckmersfound */
	uint64_t recorded;       /* s_numckmersrec */
	uint64_t dups;           /* s_ckmersasdups */
	uint64_t reads_pass;     /* s_numreadspassingjaccard */
	uint64_t reads_fail;     /* s_numreadsfailjaccard */
	uint64_t pairs_stored;   /* stored_readpairs */
	uint64_t pairs_invalid;  /* skipped_invalidreadpair */
	uint64_t pairs_nogood;   /* skipped_nogoodcontig */
} arks_oracle_map_stats;

/* ReadsProcessor::prepSeq + getStr (Common/ReadsProcessor.cpp:376-535,349-351).
 * win points at the first base of the window.  Returns 1 and fills key[ceil(k/4)]
 * or returns 0 (prepSeq's NULL). */
int arks_oracle_key(const char* win, int k, uint8_t* key);

arks_okmap* arks_oracle_kmap_new(int k, uint64_t expected_keys);
void arks_oracle_kmap_free(arks_okmap* m);
uint64_t arks_oracle_kmap_size(const arks_okmap* m);
/* returns 1 if found and sets *value */
int arks_oracle_kmap_find(const arks_okmap* m, const uint8_t* key, int32_t* value);
/* keys_out: size()*nb bytes sorted bytewise; vals_out: matching values */
void arks_oracle_kmap_dump(const arks_okmap* m, uint8_t* keys_out, int32_t* vals_out);

/* mapKmers (Arcs.cpp:869-929) for one contig end. Returns the number of k-mers added. */
int arks_oracle_map_kmers(arks_okmap* m, const char* seq, int len, int conreci, arks_oracle_index_stats* st);

/* End extraction rule of getContigKmers (Arcs.cpp:1072-1091): for a contig of
 * length len writes head=[0,cut) and tail=[len-cut,len). */
int arks_oracle_end_cutoff(int len, int end_length);

/* checkReadSequence (Arcs.cpp:366-389) */
int arks_oracle_check_read(const char* seq, int len);

/* bestContig (Arcs.cpp:939-1014) */
int arks_oracle_best_contig(const arks_okmap* m, const char* read, int len, double j_index, arks_oracle_map_stats* st);

/* The per-pair part of chromiumRead (Arcs.cpp:1266-1292) for pairs that already
 * passed the name/barcode checks: read 2i and 2i+1 are mates, read r occupies
 * bases[off[r] .. off[r+1]).  conreci_out[i] = the stored contig end or 0. */
void arks_oracle_map_pairs(const arks_okmap* m, const char* bases, const uint32_t* off, uint64_t n_pairs,
    double j_index, int32_t* conreci_out, arks_oracle_map_stats* st);

/* normalEstimation / headOrTail (Arcs.cpp:833-861); bit0 = valid, bit1 = isHead */
float arks_oracle_normal_estimation(int x, float p, int n);
int arks_oracle_head_or_tail(int head, int tail, int min_reads, float error_percent);

/* pairContigs (Arcs.cpp:1378-1435) on an imap given as rows sorted by barcode:
 * row r = (barcode[r], contig[r], head[r], tail[r]) with head+tail > 0 and one row
 * per (barcode, contig).  mult[b] is the barcode multiplicity; rank[c] the
 * std::string order of contig names.  Output: up to cap rows (a, b, counts[4]) with
 * rank[a] < rank[b], sorted by (rank[a], rank[b]).  Returns the number of rows
 * (may exceed cap; then only cap were written). */
uint64_t arks_oracle_pair_contigs(const uint32_t* barcode, const uint32_t* contig, const uint32_t* head,
    const uint32_t* tail, uint64_t n_rows, const int32_t* mult, int min_mult, int max_mult, int min_reads,
    float error_percent, const uint32_t* rank, uint32_t* out_a, uint32_t* out_b, uint32_t* out_counts,
    uint64_t cap);

/* createGraph decision (Arcs.cpp:1441-1467,1485-1498): returns 1 if the pair
 * becomes an edge and sets *orientation, *weight. */
int arks_oracle_edge(const uint32_t counts[4], int min_links, float error_percent, int* orientation, int* weight);

#ifdef __cplusplus
}
#endif
#endif
