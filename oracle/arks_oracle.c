/* oracle/arks_oracle.c -- TEST INFRASTRUCTURE ONLY (see arks_oracle.h).
 *
 * CPU restatement of the ARKS hot path of bcgsc/arcs 1.2.8.  Each function cites
 * the reference lines it follows.  Written for clarity, not speed.
 */
#include "arks_oracle.h"
#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ key -- */

/* A=0 C=1 G=2 T=3, either case; anything else -1.
 * (LUTs fw0..rv3, Common/ReadsProcessor.cpp:39-317: every other byte maps to 0xFF.) */
static int
base_code(char c)
{
	switch (c) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': return 3;
	default: return -1;
	}
}

static void
put_base(uint8_t* key, int i, int code)
{
	key[i >> 2] |= (uint8_t)(code << (6 - 2 * (i & 3)));
}

/* ReadsProcessor::prepSeq (Common/ReadsProcessor.cpp:376-535).
 *  - NULL if any base of the window is not ACGTacgt (:398-425 and the finishing loops);
 *  - forward / reverse-complement decided bytewise on the MSB-first packing, i.e.
 *    lexicographically with A<C<G<T (:427-501): the smaller one is returned;
 *  - if the window equals its reverse complement, control reaches the
 *    "palamdromic" tail (:503-534), which does NOT return the forward packing:
 *    byte h=ceil(k/8) is skipped (++outputIndex, :506), each later full byte packs
 *    4 bases but advances the cursor by only 3 (:509-518 -- index is not
 *    incremented on the fw3 lookup), and the hanging byte collapses to one base
 *    in bits 7-6 (:521-533: an unsigned char is shifted left by 2 before every
 *    OR, so only the last base written survives).  Reproduced as is. */
int
arks_oracle_key(const char* win, int k, uint8_t* key)
{
	int nb = (k + 3) / 4;
	int code[4 * ARKS_ORACLE_MAX_KEY_BYTES];
	int i;
	if (k <= 0 || k > 4 * ARKS_ORACLE_MAX_KEY_BYTES)
		return 0;
	for (i = 0; i < k; ++i) {
		code[i] = base_code(win[i]);
		if (code[i] < 0)
			return 0;
	}
	memset(key, 0, (size_t)nb);
	/* compare F with R = revcomp(F) */
	int cmp = 0;
	for (i = 0; i < k && cmp == 0; ++i) {
		int f = code[i], r = 3 - code[k - 1 - i];
		cmp = (f > r) - (f < r);
	}
	if (cmp < 0) {
		for (i = 0; i < k; ++i)
			put_base(key, i, code[i]);
		return 1;
	}
	if (cmp > 0) {
		for (i = 0; i < k; ++i)
			put_base(key, i, 3 - code[k - 1 - i]);
		return 1;
	}
	/* palindrome: deterministic garbage key */
	int h = k / 8 + (k % 8 != 0);
	int hang = k % 4;
	for (i = 0; i < 4 * h; ++i)
		put_base(key, i, code[i]);
	int idx = 4 * h;
	int o = h + 1;
	while (o + (hang ? 1 : 0) < nb) {
		key[o] = (uint8_t)((code[idx] << 6) | (code[idx + 1] << 4) | (code[idx + 2] << 2) | code[idx + 3]);
		idx += 3;
		o += 1;
	}
	if (hang) {
		/* o may equal nb for k in {6,10}: the reference writes out of bounds there
		 * (undefined); we drop the write. */
		int c = (idx < k - 1) ? code[idx + 1] : code[k - 1];
		if (o < nb)
			key[o] = (uint8_t)(c << 6);
	}
	return 1;
}

/* ----------------------------------------------------------------- kmap -- */

/* Stand-in for google::sparse_hash_map<std::string,int,CityHasher,eqstr>
 * (Arcs/Arcs.h:140-158).  Exact map keyed by the key bytes; hash and layout are
 * not result-bearing (the reference never iterates the map). */
struct arks_okmap
{
	int k, nb;
	uint64_t cap, n;
	uint8_t* keys;
	int32_t* vals;
	uint8_t* used;
};

static uint64_t
hash_bytes(const uint8_t* p, int n)
{
	uint64_t h = 1469598103934665603ull;
	for (int i = 0; i < n; ++i) {
		h ^= p[i];
		h *= 1099511628211ull;
	}
	h ^= h >> 29;
	h *= 0xbf58476d1ce4e5b9ull;
	h ^= h >> 32;
	return h;
}

static void
kmap_alloc(arks_okmap* m, uint64_t cap)
{
	m->cap = cap;
	m->keys = (uint8_t*)calloc(cap, (size_t)m->nb);
	m->vals = (int32_t*)calloc(cap, sizeof(int32_t));
	m->used = (uint8_t*)calloc(cap, 1);
}

arks_okmap*
arks_oracle_kmap_new(int k, uint64_t expected_keys)
{
	arks_okmap* m = (arks_okmap*)calloc(1, sizeof(*m));
	m->k = k;
	m->nb = (k + 3) / 4;
	uint64_t cap = 1024;
	while (cap < expected_keys * 2)
		cap <<= 1;
	kmap_alloc(m, cap);
	return m;
}

void
arks_oracle_kmap_free(arks_okmap* m)
{
	if (!m)
		return;
	free(m->keys);
	free(m->vals);
	free(m->used);
	free(m);
}

uint64_t
arks_oracle_kmap_size(const arks_okmap* m)
{
	return m->n;
}

static uint64_t
kmap_slot(const arks_okmap* m, const uint8_t* key)
{
	uint64_t s = hash_bytes(key, m->nb) & (m->cap - 1);
	while (m->used[s] && memcmp(m->keys + s * m->nb, key, (size_t)m->nb) != 0)
		s = (s + 1) & (m->cap - 1);
	return s;
}

static void
kmap_grow(arks_okmap* m)
{
	arks_okmap old = *m;
	kmap_alloc(m, old.cap * 2);
	for (uint64_t s = 0; s < old.cap; ++s)
		if (old.used[s]) {
			uint64_t t = kmap_slot(m, old.keys + s * old.nb);
			memcpy(m->keys + t * m->nb, old.keys + s * old.nb, (size_t)m->nb);
			m->vals[t] = old.vals[s];
			m->used[t] = 1;
		}
	free(old.keys);
	free(old.vals);
	free(old.used);
}

int
arks_oracle_kmap_find(const arks_okmap* m, const uint8_t* key, int32_t* value)
{
	uint64_t s = kmap_slot(m, key);
	if (!m->used[s])
		return 0;
	*value = m->vals[s];
	return 1;
}

static int g_nb_for_sort;
static int
cmp_keyidx(const void* a, const void* b)
{
	return memcmp(*(const uint8_t* const*)a, *(const uint8_t* const*)b, (size_t)g_nb_for_sort);
}

void
arks_oracle_kmap_dump(const arks_okmap* m, uint8_t* keys_out, int32_t* vals_out)
{
	const uint8_t** ptr = (const uint8_t**)malloc(sizeof(*ptr) * (m->n ? m->n : 1));
	uint64_t j = 0;
	for (uint64_t s = 0; s < m->cap; ++s)
		if (m->used[s])
			ptr[j++] = m->keys + s * m->nb;
	g_nb_for_sort = m->nb;
	qsort(ptr, j, sizeof(*ptr), cmp_keyidx);
	for (uint64_t i = 0; i < j; ++i) {
		memcpy(keys_out + i * m->nb, ptr[i], (size_t)m->nb);
		vals_out[i] = m->vals[(uint64_t)(ptr[i] - m->keys) / (uint64_t)m->nb];
	}
	free(ptr);
}

/* ---------------------------------------------------------- index build -- */

/* mapKmers (Arcs.cpp:869-929): valid window -> insert and i++; NULL window ->
 * i += k (:922-925).  Insert rule :903-920. */
int
arks_oracle_map_kmers(arks_okmap* m, const char* seq, int len, int conreci, arks_oracle_index_stats* st)
{
	uint8_t key[ARKS_ORACLE_MAX_KEY_BYTES];
	int k = m->k, num = 0, i = 0;
	if (len < k)
		return 0;
	while (i <= len - k) {
		if (arks_oracle_key(seq + i, k, key)) {
			num++;
			if ((m->n + 1) * 2 > m->cap)
				kmap_grow(m);
			uint64_t s = kmap_slot(m, key);
			if (m->used[s]) {
				if (m->vals[s] != conreci) {
					st->removed++;
					if (m->vals[s] != 0) {
						st->unique--;
						m->vals[s] = 0;
					}
				}
				st->collisions++;
			} else {
				memcpy(m->keys + s * m->nb, key, (size_t)m->nb);
				m->vals[s] = conreci;
				m->used[s] = 1;
				m->n++;
				st->unique++;
				st->recorded++;
			}
			i++;
		} else {
			i += k;
			st->kmers_null++;
		}
	}
	st->kmers_valid += (uint64_t)num;
	return num;
}

/* getContigKmers (Arcs.cpp:1072-1074) */
int
arks_oracle_end_cutoff(int len, int end_length)
{
	int cut = end_length;
	if (cut == 0 || len <= cut * 2)
		cut = len / 2;
	return cut;
}

/* --------------------------------------------------------- read mapping -- */

/* checkReadSequence (Arcs.cpp:366-389) */
int
arks_oracle_check_read(const char* seq, int len)
{
	double ambiguity = 0;
	for (int i = 0; i < len; i++) {
		char c = (char)toupper((unsigned char)seq[i]);
		if (c != 'A' && c != 'T' && c != 'G' && c != 'C') {
			if (c == 'N')
				ambiguity++;
			else
				return 0;
		}
	}
	double ar = ambiguity / (double)len;
	if (ar > 0.02)
		return 0;
	return 1;
}

typedef struct
{
	int conreci, count;
} track_t;

/* bestContig (Arcs.cpp:939-1014).  ktrack is a std::map<int,int> there: iteration
 * in increasing conreci with a strict '<' means ties go to the smallest conreci. */
int
arks_oracle_best_contig(const arks_okmap* m, const char* read, int len, double j_index, arks_oracle_map_stats* st)
{
	uint8_t key[ARKS_ORACLE_MAX_KEY_BYTES];
	int k = m->k;
	int total = 0, ntrack = 0, captrack = 16;
	track_t* track = (track_t*)malloc(sizeof(track_t) * (size_t)captrack);
	for (int i = 0; i <= len - k; ++i) {
		total++;
		if (arks_oracle_key(read + i, k, key)) {
			int32_t v;
			st->kmers_valid++;
			if (arks_oracle_kmap_find(m, key, &v)) {
				if (v != 0) {
					int t = 0;
					while (t < ntrack && track[t].conreci != v)
						t++;
					if (t == ntrack) {
						if (ntrack == captrack) {
							captrack *= 2;
							track = (track_t*)realloc(track, sizeof(track_t) * (size_t)captrack);
						}
						track[ntrack].conreci = v;
						track[ntrack].count = 0;
						ntrack++;
					}
					track[t].count++;
					st->recorded++;
				} else {
					st->dups++;
				}
				st->found++;
			}
		} else {
			st->kmers_invalid++;
		}
	}
	double maxj = 0;
	int best = 0;
	/* emulate ordered-map iteration: visit conrecis in increasing order */
	for (int a = 0; a < ntrack; ++a)
		for (int b = a + 1; b < ntrack; ++b)
			if (track[b].conreci < track[a].conreci) {
				track_t t = track[a];
				track[a] = track[b];
				track[b] = t;
			}
	for (int t = 0; t < ntrack; ++t) {
		double cur = (double)track[t].count / (double)total;
		if (maxj < cur) {
			maxj = cur;
			best = track[t].conreci;
		}
	}
	free(track);
	if (maxj > j_index) {
		st->reads_pass++;
		return best;
	}
	st->reads_fail++;
	return 0;
}

/* chromiumRead, per-pair body after the name/barcode checks (Arcs.cpp:1266-1292) */
void
arks_oracle_map_pairs(const arks_okmap* m, const char* bases, const uint32_t* off, uint64_t n_pairs,
    double j_index, int32_t* conreci_out, arks_oracle_map_stats* st)
{
	for (uint64_t p = 0; p < n_pairs; ++p) {
		const char* r1 = bases + off[2 * p];
		const char* r2 = bases + off[2 * p + 1];
		int l1 = (int)(off[2 * p + 1] - off[2 * p]);
		int l2 = (int)(off[2 * p + 2] - off[2 * p + 1]);
		int c1 = 0, c2 = 0;
		if (arks_oracle_check_read(r1, l1) && arks_oracle_check_read(r2, l2)) {
			c1 = arks_oracle_best_contig(m, r1, l1, j_index, st);
			c2 = arks_oracle_best_contig(m, r2, l2, j_index, st);
		} else {
			st->pairs_invalid++;
		}
		if (c1 != 0 && c1 == c2) {
			conreci_out[p] = c1;
			st->pairs_stored++;
		} else {
			conreci_out[p] = 0;
			st->pairs_nogood++;
		}
	}
}

/* ----------------------------------------------------------- pair links -- */

/* normalEstimation (Arcs.cpp:833-839).  Types matter: mean and sd are float,
 * std::sqrt(2) is the double overload, std::erf is evaluated in double and the
 * result is narrowed to float on return. */
float
arks_oracle_normal_estimation(int x, float p, int n)
{
	float mean = n * p;
	float sd = sqrtf(n * p * (1 - p));
	return (float)(0.5 * (1 + erf((x - mean) / (sd * sqrt(2.0)))));
}

/* headOrTail (Arcs.cpp:846-861): ties between head and tail go to head. */
int
arks_oracle_head_or_tail(int head, int tail, int min_reads, float error_percent)
{
	int max = head > tail ? head : tail;
	int sum = head + tail;
	if (sum < min_reads)
		return 0;
	float cdf = arks_oracle_normal_estimation(max, 0.5f, sum);
	if (1 - cdf < error_percent)
		return 1 | ((max == head) ? 2 : 0);
	return 0;
}

typedef struct
{
	uint32_t ra, rb, a, b, orient;
} link_t;

static int
cmp_link(const void* x, const void* y)
{
	const link_t* p = (const link_t*)x;
	const link_t* q = (const link_t*)y;
	if (p->ra != q->ra)
		return p->ra < q->ra ? -1 : 1;
	if (p->rb != q->rb)
		return p->rb < q->rb ? -1 : 1;
	return 0;
}

/* pairContigs (Arcs.cpp:1378-1435) */
uint64_t
arks_oracle_pair_contigs(const uint32_t* barcode, const uint32_t* contig, const uint32_t* head,
    const uint32_t* tail, uint64_t n_rows, const int32_t* mult, int min_mult, int max_mult, int min_reads,
    float error_percent, const uint32_t* rank, uint32_t* out_a, uint32_t* out_b, uint32_t* out_counts,
    uint64_t cap)
{
	uint64_t nl = 0, capl = 1024;
	link_t* links = (link_t*)malloc(sizeof(link_t) * capl);
	uint64_t r0 = 0;
	while (r0 < n_rows) {
		uint64_t r1 = r0;
		while (r1 < n_rows && barcode[r1] == barcode[r0])
			r1++;
		int mu = mult[barcode[r0]];
		if (mu >= min_mult && mu <= max_mult) {
			for (uint64_t o = r0; o < r1; ++o)
				for (uint64_t p = r0; p < r1; ++p) {
					if (!(rank[contig[o]] < rank[contig[p]]))
						continue;
					int ha = arks_oracle_head_or_tail((int)head[o], (int)tail[o], min_reads, error_percent);
					int hb = arks_oracle_head_or_tail((int)head[p], (int)tail[p], min_reads, error_percent);
					if ((ha & 1) && (hb & 1)) {
						if (nl == capl) {
							capl *= 2;
							links = (link_t*)realloc(links, sizeof(link_t) * capl);
						}
						links[nl].a = contig[o];
						links[nl].b = contig[p];
						links[nl].ra = rank[contig[o]];
						links[nl].rb = rank[contig[p]];
						links[nl].orient = (uint32_t)(((ha & 2) ? 0 : 2) + ((hb & 2) ? 0 : 1));
						nl++;
					}
				}
		}
		r0 = r1;
	}
	qsort(links, nl, sizeof(link_t), cmp_link);
	uint64_t n_out = 0;
	for (uint64_t i = 0; i < nl;) {
		uint64_t j = i;
		uint32_t c[4] = { 0, 0, 0, 0 };
		while (j < nl && links[j].ra == links[i].ra && links[j].rb == links[i].rb)
			c[links[j++].orient]++;
		if (n_out < cap) {
			out_a[n_out] = links[i].a;
			out_b[n_out] = links[i].b;
			memcpy(out_counts + 4 * n_out, c, sizeof(c));
		}
		n_out++;
		i = j;
	}
	free(links);
	return n_out;
}

/* getMaxValueAndIndex + the "second" rule of createGraph + checkSignificance
 * (Arcs.cpp:1441-1467,1485-1498) */
int
arks_oracle_edge(const uint32_t counts[4], int min_links, float error_percent, int* orientation, int* weight)
{
	unsigned max = 0, index = 0, second = 0;
	for (unsigned i = 0; i < 4; ++i)
		if (counts[i] > max) {
			max = counts[i];
			index = i;
		}
	for (unsigned i = 0; i < 4; ++i)
		if (counts[i] != max && counts[i] > second)
			second = counts[i];
	*orientation = (int)index;
	*weight = (int)max;
	if ((int)max < min_links)
		return 0;
	float cdf = arks_oracle_normal_estimation((int)max, 0.5f, (int)(max + second));
	return (1 - cdf < error_percent) ? 1 : 0;
}
