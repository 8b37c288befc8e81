"""Seeded synthetic drafts + barcoded linked reads for the ARKS hot path.

Draft: i.i.d. ACGT with planted features the reference's semantics are sensitive to --
N runs (lengths around k), sequence duplicated across contigs (value-0 keys), (AT)n /
(CG)n / (ACGT)n tracts (windows equal to their own reverse complement), lower case,
IUPAC codes.  Reads: molecules sampled from the true genome order so neighbouring
contigs share barcodes; 2 x read_len pairs (mate 2 reverse-complemented), substitution
errors and Ns.  numpy only; everything derives from the seed.
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGTacgtNn", b"TGCAtgcaNn"):
    _COMP[a] = b


def revcomp(a):
    return _COMP[a[::-1]]


def make_draft(rng, genome_len, mean_contig, k, n_runs=4, dup_frac=0.01, palindromes=3, lower_frac=0.05,
               iupac=2, min_contig=600):
    """-> (genome uint8, contigs list of (name, start, end)) ; contig i covers genome[start:end]"""
    g = ACGT[rng.integers(0, 4, genome_len)].copy()
    # duplicated segments (copied elsewhere in the genome -> keys shared by several contig ends)
    n_dup = max(1, int(genome_len * dup_frac / (4 * k)))
    for _ in range(n_dup):
        L = int(rng.integers(k, 4 * k))
        a, b = rng.integers(0, genome_len - L, 2)
        g[b:b + L] = g[a:a + L]
    # palindromic tracts
    for i in range(palindromes):
        unit = [b"AT", b"CG", b"ACGT"][i % 3]
        L = 2 * k + int(rng.integers(0, k))
        tract = np.frombuffer((unit * (L // len(unit) + 1))[:L], dtype=np.uint8)
        p = int(rng.integers(0, genome_len - L))
        g[p:p + L] = tract
    # N runs of assorted lengths (1 .. 3k), singles, and closely spaced pairs
    for _ in range(n_runs):
        L = int(rng.integers(1, 3 * k))
        p = int(rng.integers(0, genome_len - L))
        g[p:p + L] = ord("N")
        if rng.random() < 0.5:
            q = min(genome_len - 1, p + L + int(rng.integers(1, k)))
            g[q] = ord("n")
    for _ in range(iupac):
        g[int(rng.integers(0, genome_len))] = rng.choice(np.frombuffer(b"RYKMSWryN", dtype=np.uint8))
    lower = rng.random(genome_len) < lower_frac
    g = np.where(lower, g | 0x20, g).astype(np.uint8)
    # cut into contigs (log-normal lengths) separated by small gaps
    contigs, pos, i = [], 0, 1
    while pos < genome_len:
        L = int(max(min_contig // 2, rng.lognormal(np.log(mean_contig), 0.5)))
        end = min(genome_len, pos + L)
        contigs.append((str(i), pos, end))
        i += 1
        pos = end + int(rng.integers(0, 50))
    return g, contigs


def make_reads(rng, genome, n_barcodes, pairs_per_barcode, read_len=150, mol_len=50000, mols_per_barcode=4,
               insert=350, sub_rate=0.002, n_rate=0.001, len_jitter=0):
    """-> (bases uint8, read_off uint32[2n+1], barcode_id uint32[n]) grouped by barcode"""
    G = len(genome)
    n = n_barcodes * pairs_per_barcode
    bc = np.repeat(np.arange(n_barcodes, dtype=np.uint32), pairs_per_barcode)
    mol_start = rng.integers(0, max(1, G - mol_len), (n_barcodes, mols_per_barcode))
    which = rng.integers(0, mols_per_barcode, n)
    ms = mol_start[bc, which]
    span = max(1, mol_len - insert - read_len)
    p1 = np.minimum(ms + rng.integers(0, span, n), G - insert - read_len - 1)
    p2 = p1 + insert
    if len_jitter:
        l1 = read_len - rng.integers(0, len_jitter + 1, n)
        l2 = read_len - rng.integers(0, len_jitter + 1, n)
    else:
        l1 = np.full(n, read_len)
        l2 = np.full(n, read_len)
    lens = np.empty(2 * n, dtype=np.int64)
    lens[0::2], lens[1::2] = l1, l2
    off = np.zeros(2 * n + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    bases = np.empty(off[-1], dtype=np.uint8)
    flip = rng.random(n) < 0.5  # molecule strand
    for i in range(n):
        a = genome[p1[i]:p1[i] + l1[i]]
        b = revcomp(genome[p2[i] + read_len - l2[i]:p2[i] + read_len])
        if flip[i]:
            a, b = b, a
            l1[i], l2[i] = l2[i], l1[i]
        bases[off[2 * i]:off[2 * i] + len(a)] = a
        bases[off[2 * i] + len(a):off[2 * i + 2]] = b
        off[2 * i + 1] = off[2 * i] + len(a)
    sub = rng.random(len(bases)) < sub_rate
    bases[sub] = ACGT[rng.integers(0, 4, int(sub.sum()))]
    nn = rng.random(len(bases)) < n_rate
    bases[nn] = ord("N")
    return bases, off.astype(np.uint32), bc


def contig_end_arrays(genome, contigs, k, min_size=500, end_length=30000):
    """getContigKmers' end extraction (Arcs.cpp:1056-1091) -> (bases, end_off uint64, conreci uint32, names)"""
    chunks, lens, conreci, names = [], [], [], []
    for name, s, e in contigs:
        L = e - s
        if L < min_size:
            continue
        i = len(names)
        names.append(name)
        cut = end_length
        if cut == 0 or L <= 2 * cut:
            cut = L // 2
        chunks += [genome[s:s + cut], genome[e - cut:e]]
        lens += [cut, cut]
        conreci += [2 * i + 1, 2 * i + 2]
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    bases = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint8)
    return bases, off, np.array(conreci, dtype=np.uint32), names
