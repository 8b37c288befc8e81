"""GPU parity: the CUDA path (through the C ABI of libarks_b200.so) against the CPU
restatement (oracle/) on the same seeded inputs -- bit-exact for every integer result --
and against the reference's golden demo outputs."""
import os

import numpy as np
import pytest

import glue
import oracle_lib as O
import seqio
from tools import synth

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _arks():
    import arcs_b200
    return arcs_b200


def _sorted_rows(keys, vals):
    if len(vals) == 0:
        return keys, vals
    order = np.lexsort(keys.T[::-1])
    return keys[order], vals[order]


def _oracle_index(k, bases, end_off, conreci):
    km = O.KMap(k, int(end_off[-1]))
    for e in range(len(conreci)):
        km.map_kmers(bases[int(end_off[e]):int(end_off[e + 1])].tobytes(), int(conreci[e]))
    return km


def _check_index(k, bases, end_off, conreci, split=None):
    A = _arks()
    km = _oracle_index(k, bases, end_off, conreci)
    idx = A.ArksIndex(k, int(end_off[-1]) + 16)
    if split:  # several add calls
        for a in range(0, len(conreci), split):
            b = min(len(conreci), a + split)
            idx.add_ends(bases, end_off[a:b + 1], conreci[a:b])
    else:
        idx.add_ends(bases, end_off, conreci)
    st = idx.finalize().as_dict()
    ref = km.stats.as_dict()
    assert st == ref, (st, ref)
    gk, gv = _sorted_rows(*idx.dump())
    ok, ov = km.dump()
    assert gk.shape == ok.shape
    assert np.array_equal(gk, ok)
    assert np.array_equal(gv, ov)
    return idx, km


@pytest.mark.parametrize("k", [4, 5, 8, 16, 20, 21, 30, 31, 32, 33, 40, 47, 60, 63, 64])
def test_index_build_matches_oracle(k):
    rng = np.random.default_rng(100 + k)
    genome, contigs = synth.make_draft(rng, 60000, 4000, k, n_runs=12, palindromes=6, iupac=6)
    bases, end_off, conreci, _ = synth.contig_end_arrays(genome, contigs, k, min_size=500, end_length=1500)
    _check_index(k, bases, end_off, conreci, split=None if k % 2 else 5)


def test_index_edge_cases():
    k = 20
    seqs = [
        b"ACGT" * 3,  # shorter than k: ignored
        b"A" * 19,
        b"ACGTTGCATGCATGCATGCA",  # exactly one window
        b"N" * 50,
        b"ACGTACGTAC" * 5 + b"N" + b"TTGACCAGTA" * 5,
        b"GATTACAGATTACAGATTACAN",  # N in the last position
        b"NGATTACAGATTACAGATTACA",
        b"AT" * 40,  # palindromes everywhere
        b"CG" * 40,
        b"acgtnACGT" * 20,
        b"ACGTTGCATGCATGCATGCA",  # duplicate of an earlier end -> value 0
        b"TGCATGCATGCATGCAACGT",  # its reverse complement
        b"ACGTTGCATGCATGCATGCANNACGTTGCATGCATGCATGCATTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT",  # NN: jump skips valid windows
        b"",
    ]
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    end_off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    end_off[1:] = np.cumsum([len(s) for s in seqs])
    conreci = np.arange(1, len(seqs) + 1, dtype=np.uint32)
    _check_index(k, bases, end_off, conreci)


def _check_map(k, j, genome, contigs, rng, **read_kw):
    bases, end_off, conreci, names = synth.contig_end_arrays(genome, contigs, k, min_size=500, end_length=3000)
    idx, km = _check_index(k, bases, end_off, conreci)
    rb, roff, bc = synth.make_reads(rng, genome, **read_kw)
    got = idx.map_pairs(rb, roff, bc, j)
    want, st = km.map_pairs(rb, roff, j)
    assert np.array_equal(got, want), np.nonzero(got != want)[0][:10]
    assert idx.map_stats().as_dict() == st.as_dict()
    return idx, km, names, bc, got


@pytest.mark.parametrize("k,j", [(20, 0.05), (30, 0.55), (32, 0.3), (33, 0.3), (40, 0.5), (60, 0.55), (64, 0.2)])
def test_map_pairs_matches_oracle(k, j):
    rng = np.random.default_rng(7 * k)
    genome, contigs = synth.make_draft(rng, 120000, 8000, k, n_runs=10, palindromes=4)
    _check_map(k, j, genome, contigs, rng, n_barcodes=40, pairs_per_barcode=30, read_len=150, mol_len=20000,
               mols_per_barcode=2, sub_rate=0.004, n_rate=0.002, len_jitter=40)


def test_map_pairs_ragged_and_long_reads():
    k, j = 24, 0.1
    rng = np.random.default_rng(5)
    genome, contigs = synth.make_draft(rng, 80000, 9000, k)
    bases, end_off, conreci, _ = synth.contig_end_arrays(genome, contigs, k, end_length=4000)
    idx, km = _check_index(k, bases, end_off, conreci)
    reads = []
    for L in [0, 1, k - 1, k, k + 1, 100, 511, 512, 513, 700, 1024, 2500, 150, 150]:
        p = int(rng.integers(0, len(genome) - L - 1))
        r = genome[p:p + L].copy()
        reads.append(r)
    reads.append(np.frombuffer(b"ACGTNNNNNNNNACGT" * 10, dtype=np.uint8))  # too many N
    reads.append(genome[100:250].copy())
    bad = genome[300:450].copy()
    bad[7] = ord("R")  # non-ACGTN character: invalid pair
    reads.append(bad)
    reads.append(genome[300:450].copy())
    assert len(reads) % 2 == 0
    rb = np.concatenate(reads)
    roff = np.zeros(len(reads) + 1, dtype=np.uint32)
    roff[1:] = np.cumsum([len(r) for r in reads])
    bc = np.arange(len(reads) // 2, dtype=np.uint32)
    got = idx.map_pairs(rb, roff, bc, j)
    want, st = km.map_pairs(rb, roff, j)
    assert np.array_equal(got, want)
    assert idx.map_stats().as_dict() == st.as_dict()


@pytest.mark.parametrize("k,j", [(20, 0.05), (31, 0.2), (60, 0.3)])
def test_map_pairs_chimeric_noisy_and_boundary_reads(k, j):
    """reads that leave the lane-per-read path of map_groups_kernel in every way it has: votes for a
    second contig end (chimeras, contig-boundary reads), more staged lookups per group than one round
    holds (dense errors in 250-bp reads), N runs at both ends, lower case, both strands"""
    rng = np.random.default_rng(1000 + k)
    genome, contigs = synth.make_draft(rng, 150000, 6000, k, n_runs=6, palindromes=3)
    bases, end_off, conreci, _ = synth.contig_end_arrays(genome, contigs, k, end_length=2500)
    idx, km = _check_index(k, bases, end_off, conreci)
    G = len(genome)
    reads = []

    def piece(L):
        p = int(rng.integers(0, G - L - 1))
        r = genome[p:p + L].copy()
        return synth.revcomp(r) if rng.random() < 0.5 else r

    for i in range(400):
        kind = i % 8
        L = int(rng.integers(max(k, 60), 257))
        if kind == 0:  # chimera of two loci
            a = int(rng.integers(k // 2, L - k // 2 + 1)) if L > k else L
            r = np.concatenate([piece(a), piece(L - a)]) if 0 < a < L else piece(L)
        elif kind == 1:  # dense substitutions: most windows have to be looked up
            r = piece(L)
            pos = np.arange(int(rng.integers(0, k)), L, max(2, k // 2 + int(rng.integers(0, k))))
            r[pos] = synth.ACGT[rng.integers(0, 4, len(pos))]
        elif kind == 2:  # Ns at the ends and in the middle (at most 2 %)
            r = piece(L)
            for q in (0, L - 1, L // 2)[: 1 + int(rng.integers(0, 3))]:
                r[q] = ord("N") if rng.random() < 0.7 else ord("n")
        elif kind == 3:  # straddles a contig boundary
            name, s0, e0 = contigs[int(rng.integers(0, len(contigs) - 1))]
            p = max(0, min(G - L - 1, e0 - int(rng.integers(1, L))))
            r = genome[p:p + L].copy()
        elif kind == 4:  # one error near each end: the first two seeds miss
            r = piece(L)
            r[int(rng.integers(0, min(k, L)))] = ord("A")
            r[L - 1 - int(rng.integers(0, min(k, L)))] = ord("C")
        elif kind == 5:  # lower case + a single substitution
            r = piece(L) | 0x20
            r[int(rng.integers(0, L))] = ord("g")
        elif kind == 6:  # random sequence: nothing found
            r = synth.ACGT[rng.integers(0, 4, L)].copy()
        else:
            r = piece(L)
        reads.append(r.astype(np.uint8))
    rb = np.concatenate(reads)
    roff = np.zeros(len(reads) + 1, dtype=np.uint32)
    roff[1:] = np.cumsum([len(r) for r in reads])
    bc = (np.arange(len(reads) // 2) // 10).astype(np.uint32)
    got = idx.map_pairs(rb, roff, bc, j)
    want, st = km.map_pairs(rb, roff, j)
    assert np.array_equal(got, want), np.nonzero(got != want)[0][:10]
    assert idx.map_stats().as_dict() == st.as_dict()


def test_map_pairs_dense_table(monkeypatch):
    """load factor 0.9 (what a draft that nearly fills the GPU gets): long probe sequences in the build, in
    the seed probes and behind the membership filter"""
    monkeypatch.setenv("ARKS_TABLE_LOAD", "0.9")
    k, j = 40, 0.3
    rng = np.random.default_rng(99)
    genome, contigs = synth.make_draft(rng, 100000, 7000, k, n_runs=3, dup_frac=0.0)
    _check_map(k, j, genome, contigs, rng, n_barcodes=30, pairs_per_barcode=30, read_len=150, mol_len=20000,
               mols_per_barcode=2, sub_rate=0.004, n_rate=0.002, len_jitter=20)


@pytest.mark.parametrize("k,piece,n_pieces", [(10, 12, 40), (20, 25, 40), (20, 25, 150), (16, 17, 300)])
def test_reads_voting_for_more_than_32_contig_ends(k, piece, n_pieces):
    """bestContig keeps its votes in an unbounded std::map (Arcs.cpp:951-1004); the device tracks 32 contig ends per
    warp and, for a read that votes for more, counts in 2, 4, ... passes over its windows.  Reads stitched from
    pieces of many contigs: up to 512 bases (one packed region) and longer (the chunked path), every piece one or
    two windows long so that many ends tie; the winner gets an extra piece in some reads."""
    rng = np.random.default_rng(k * 1000 + n_pieces)
    n_contigs = max(64, n_pieces)
    contigs = [(str(i + 1), synth.ACGT[rng.integers(0, 4, 700)].tobytes()) for i in range(n_contigs)]
    ends, names = glue.contig_ends(contigs, 500, 300)
    bases = np.frombuffer(b"".join(s for s, _ in ends), dtype=np.uint8)
    end_off = np.zeros(len(ends) + 1, dtype=np.uint64)
    end_off[1:] = np.cumsum([len(s) for s, _ in ends])
    conreci = np.array([cr for _, cr in ends], dtype=np.uint32)
    idx, km = _check_index(k, bases, end_off, conreci)
    reads = []
    for r in range(24):
        order = rng.permutation(len(ends))[:n_pieces]
        pieces = []
        for e in order:
            seq = ends[e][0]
            a = int(rng.integers(0, len(seq) - piece))
            pieces.append(seq[a:a + piece])
        if r % 3 == 0:  # one end gets a second piece: a unique winner
            seq = ends[order[n_pieces // 2]][0]
            pieces.append(seq[5:5 + piece])
        read = b"".join(pieces)
        mate = synth.revcomp(np.frombuffer(read, dtype=np.uint8)).tobytes() if r % 2 else read
        reads += [read, mate]
    rb = np.frombuffer(b"".join(reads), dtype=np.uint8)
    roff = np.zeros(len(reads) + 1, dtype=np.uint32)
    roff[1:] = np.cumsum([len(x) for x in reads])
    bc = np.arange(len(reads) // 2, dtype=np.uint32)
    stored = 0
    for j in (0.0, 0.004, 0.02):
        idx.map_stats_reset()
        got = idx.map_pairs(rb, roff, bc, j)
        want, st = km.map_pairs(rb, roff, j)
        assert np.array_equal(got, want), (j, got, want)
        assert idx.map_stats().as_dict() == st.as_dict()
        stored += int((want != 0).sum())
    assert stored > 0  # some of these pairs do get a contig end


def _decode_key(key_bytes, k):
    return bytes(b"ACGT"[(key_bytes[i // 4] >> (6 - 2 * (i % 4))) & 3] for i in range(k))


@pytest.mark.parametrize("k", [20, 30, 32, 40, 60, 64])
def test_substitutions_around_palindromes_and_garbage_keys(k):
    """Exactness corners of any shortcut around a mismatching base (seed-and-extend, membership filters): palindromic
    read windows (looked up under the reference's garbage key, ReadsProcessor.cpp:503-534), genuine windows that equal
    the garbage key of a palindromic text window, errors at the very ends of a read / next to each other, and errors
    inside palindromic tracts."""
    rng = np.random.default_rng(900 + k)
    rnd = lambda n: synth.ACGT[rng.integers(0, 4, n)].tobytes()  # noqa: E731
    tracts = [b"AT" * 100, b"CG" * 100, b"ACGT" * 60, b"AATT" * 60, b"GAATTC" * 40]
    pal = tracts[0][:k]
    garbage = _decode_key(O.key(pal, k), k)  # the k-mer the garbage key of an (AT)n window spells
    assert O.key(garbage, k) is not None
    contigs = [("c%d" % i, rnd(700) + t + rnd(700)) for i, t in enumerate(tracts)]
    contigs.append(("g", rnd(600) + garbage + rnd(600)))  # ... present in the draft as a genuine k-mer
    contigs += [("r%d" % i, rnd(1500)) for i in range(4)]
    ends, names = glue.contig_ends(contigs, 500, 30000)
    bases = np.frombuffer(b"".join(s for s, _ in ends), dtype=np.uint8)
    end_off = np.zeros(len(ends) + 1, dtype=np.uint64)
    end_off[1:] = np.cumsum([len(s) for s, _ in ends])
    conreci = np.array([cr for _, cr in ends], dtype=np.uint32)
    idx, km = _check_index(k, bases, end_off, conreci)
    reads = []
    L = 150 if k <= 60 else 200
    for name, seq in contigs:
        for _ in range(40):
            a = int(rng.integers(0, len(seq) - L))
            r = bytearray(seq[a:a + L])
            for _ in range(int(rng.integers(1, 4))):  # 1-3 substitutions, some at the very ends, some adjacent
                p = int(rng.choice([0, 1, L - 1, L - 2, int(rng.integers(0, L)), int(rng.integers(0, L))]))
                r[p] = b"ACGT"[(b"ACGT".index(bytes([r[p]]).upper()) + int(rng.integers(1, 4))) % 4] if bytes([r[p]]).upper() in b"ACGT" else r[p]
                if rng.random() < 0.3 and p + 1 < L:
                    r[p + 1] = ord("A") if r[p + 1] != ord("A") else ord("C")
            reads.append(bytes(r))
    # reads that carry a palindromic window next to an error, and the garbage k-mer next to an error
    for t in tracts:
        core = t[:k + 20]
        reads.append(rnd(40) + core + rnd(L - 40 - len(core)) if L > 40 + len(core) else core[:L])
    reads.append(rnd(30) + garbage + rnd(L - 30 - k))
    if len(reads) % 2:
        reads.append(reads[-1])
    mates = []
    for i, r in enumerate(reads):
        mates += [r, synth.revcomp(np.frombuffer(r, dtype=np.uint8)).tobytes() if i % 2 else r]
    rb = np.frombuffer(b"".join(mates), dtype=np.uint8)
    roff = np.zeros(len(mates) + 1, dtype=np.uint32)
    roff[1:] = np.cumsum([len(x) for x in mates])
    bc = np.arange(len(mates) // 2, dtype=np.uint32)
    for j in (0.05, 0.3):
        idx.map_stats_reset()
        got = idx.map_pairs(rb, roff, bc, j)
        want, st = km.map_pairs(rb, roff, j)
        assert np.array_equal(got, want), (j, np.nonzero(got != want)[0][:10])
        assert idx.map_stats().as_dict() == st.as_dict()


def test_pair_links_match_oracle():
    k, j = 30, 0.4
    rng = np.random.default_rng(11)
    genome, contigs = synth.make_draft(rng, 300000, 6000, k)
    idx, km, names, bc, conreci = _check_map(k, j, genome, contigs, rng, n_barcodes=150, pairs_per_barcode=60,
                                            mol_len=30000, mols_per_barcode=3)
    # imap parity
    barcodes = [str(b) for b in bc]
    imap, uniq = glue.imap_rows(barcodes, conreci, names)
    gb, gc, gh, gt = idx.imap()
    got_rows = sorted(zip(gb.tolist(), gc.tolist(), gh.tolist(), gt.tolist()))
    want_rows = sorted((int(b), c, ht[0], ht[1]) for b, d in imap.items() for c, ht in d.items())
    assert got_rows == want_rows
    # pmap parity for a few parameter sets
    rank = glue.lex_rank(uniq)
    nb = int(bc.max()) + 1
    mult = np.bincount(bc, minlength=nb).astype(np.int32) * 2
    rows = np.array(want_rows, dtype=np.uint32)
    for (c, r, mlo, mhi) in [(5, 0.05, 50, 10000), (2, 0.2, 1, 100000), (3, 0.01, 120, 121), (1, 0.5, 0, 10 ** 6)]:
        a, b, counts = idx.pair_links(mult, mlo, mhi, c, r, rank)
        oa, ob, oc = O.pair_contigs(rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3], mult, mlo, mhi, c, np.float32(r), rank)
        assert np.array_equal(a, oa) and np.array_equal(b, ob) and np.array_equal(counts, oc)
        assert glue.gv_text(a, b, counts, uniq, rank, 0, np.float32(r)) == glue.gv_text(oa, ob, oc, uniq, rank, 0, np.float32(r))


def _gpu_demo(fa, fq, k, j, c, m, e, z, r, l, multfile=None):
    A = _arks()
    contigs = seqio.read_fasta(fa)
    ends, names = glue.contig_ends(contigs, z, e)
    bases = np.frombuffer(b"".join(s for s, _ in ends), dtype=np.uint8)
    end_off = np.zeros(len(ends) + 1, dtype=np.uint64)
    end_off[1:] = np.cumsum([len(s) for s, _ in ends])
    conreci = np.array([cr for _, cr in ends], dtype=np.uint32)
    idx = A.ArksIndex(k, int(end_off[-1]))
    idx.add_ends(bases, end_off, conreci)
    ist = idx.finalize().as_dict()
    recs = seqio.read_fastq(fq)
    if multfile:
        mult = {ln.split()[0]: int(ln.split()[1]) for ln in open(multfile)}
    else:
        mult = seqio.multiplicities(recs)
    barcodes, rb, roff = seqio.candidate_pairs(recs, mult)
    bnames = sorted(set(barcodes))
    bid = {b: i for i, b in enumerate(bnames)}
    bc = np.array([bid[b] for b in barcodes], dtype=np.uint32)
    # feed in several batches, as the host pipeline does
    n = len(bc)
    step = max(1, n // 3)
    for a in range(0, n, step):
        b = min(n, a + step)
        idx.map_pairs(rb, roff[2 * a:2 * b + 1], bc[a:b], j, want_conreci=False)
    mst = idx.map_stats().as_dict()
    ids, uniq = glue.name_ids(names)
    rank = glue.lex_rank(uniq)
    multv = np.array([mult.get(b, 0) for b in bnames], dtype=np.int32)
    a, b, counts = idx.pair_links(multv, m[0], m[1], c, r, rank)
    return ist, mst, glue.gv_text(a, b, counts, uniq, rank, l, np.float32(r)), idx


def test_arks_demo_golden_gpu():
    d = os.path.join(GOLD, "arks_demo")
    ist, mst, gv, idx = _gpu_demo(os.path.join(d, "test_scaffolds.renamed.fa"), os.path.join(d, "test_reads.fq.gz"),
                                  k=30, j=0.55, c=5, m=(50, 6000), e=30000, z=500, r=0.05, l=0)
    assert gv == open(os.path.join(d, "expected_original.gv")).read()
    assert (ist["kmers_valid"], ist["kmers_null"], ist["recorded"], ist["collisions"], ist["removed"], ist["unique"]) == (
        123190, 303, 118710, 4480, 547, 118334)
    assert (mst["pairs_stored"], mst["pairs_invalid"], mst["pairs_nogood"]) == (21632, 0, 6012)
    assert (mst["kmers_valid"], mst["kmers_invalid"], mst["found"], mst["recorded"], mst["dups"]) == (
        6109324, 0, 4862376, 4814099, 48277)
    assert (mst["reads_pass"], mst["reads_fail"]) == (44503, 10785)
    assert idx.launches > 0


def test_arks_long_demo_golden_gpu():
    d = os.path.join(GOLD, "arks_long_demo")
    _, _, gv, _ = _gpu_demo(os.path.join(d, "test_scaffolds.renamed.fa"), os.path.join(d, "test_reads.cut250.fq.gz"),
                            k=20, j=0.05, c=3, m=(8, 10000), e=30000, z=500, r=0.05, l=0,
                            multfile=os.path.join(d, "barcodeMultiplicityArcs.tsv"))
    assert gv == open(os.path.join(d, "expected_original.gv")).read()


def test_device_pointer_entry_points_match_host_ones():
    """arks_index_add_device / arks_map_pairs_device (what bench.py's device-resident leg calls)"""
    import torch
    A = _arks()
    k, j = 40, 0.5
    rng = np.random.default_rng(21)
    genome, contigs = synth.make_draft(rng, 150000, 9000, k)
    bases, end_off, conreci, _ = synth.contig_end_arrays(genome, contigs, k, end_length=4000)
    rb, roff, bc = synth.make_reads(rng, genome, n_barcodes=30, pairs_per_barcode=40, mol_len=20000, mols_per_barcode=2, len_jitter=20)
    host = A.ArksIndex(k, int(end_off[-1]))
    host.add_ends(bases, end_off, conreci)
    hst = host.finalize().as_dict()
    want = host.map_pairs(rb, roff, bc, j)
    dev = A.ArksIndex(k, int(end_off[-1]))
    stream = torch.cuda.current_stream()
    dev.set_stream(stream.cuda_stream)
    d_bases = torch.from_numpy(bases.copy()).cuda()
    d_off = torch.from_numpy(end_off.astype(np.int64)).cuda()
    d_cr = torch.from_numpy(conreci.astype(np.int32)).cuda()
    dev.add_ends_device(d_bases.data_ptr(), d_off.data_ptr(), d_cr.data_ptr(), end_off)
    assert dev.finalize().as_dict() == hst
    d_rb = torch.from_numpy(rb.copy()).cuda()
    d_roff = torch.from_numpy(roff.astype(np.int64)).to(torch.int32).cuda()
    d_bc = torch.from_numpy(bc.astype(np.int32)).cuda()
    d_out = torch.zeros(len(bc), dtype=torch.int32, device="cuda")
    dev.map_pairs_device(d_rb.data_ptr(), d_roff.data_ptr(), d_bc.data_ptr(), len(bc), len(rb), j, d_out.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), want)
    assert dev.map_stats().as_dict() == host.map_stats().as_dict()
    hi, di = host.imap(), dev.imap()
    assert sorted(zip(*[x.tolist() for x in hi])) == sorted(zip(*[x.tolist() for x in di]))


def test_empty_and_degenerate_inputs():
    A = _arks()
    k = 30
    idx = A.ArksIndex(k, 1000)
    # ends: empty, shorter than k, all N
    seqs = [b"", b"ACGT", b"N" * 100]
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    idx.add_ends(bases, off, np.arange(1, len(seqs) + 1, dtype=np.uint32))
    st = idx.finalize().as_dict()
    assert st["recorded"] == 0 and st["kmers_valid"] == 0 and st["kmers_null"] == 3  # 71 NULL windows visited every k
    keys, vals = idx.dump()
    assert len(vals) == 0
    # zero pairs, then pairs of empty / tiny / all-N reads
    idx.map_pairs(np.zeros(0, np.uint8), np.zeros(1, np.uint32), np.zeros(0, np.uint32), 0.5)
    reads = [b"", b"", b"ACG", b"ACGTACGT", b"N" * 60, b"A" * 60, b"A" * 60, b"C" * 60]
    rb = np.frombuffer(b"".join(reads), dtype=np.uint8)
    roff = np.zeros(len(reads) + 1, dtype=np.uint32)
    roff[1:] = np.cumsum([len(r) for r in reads])
    got = idx.map_pairs(rb, roff, np.arange(4, dtype=np.uint32), 0.5)
    km = O.KMap(k, 16)
    want, ost = km.map_pairs(rb, roff, 0.5)
    assert np.array_equal(got, want) and not got.any()
    assert idx.map_stats().as_dict() == ost.as_dict()
    a, b, c = idx.pair_links(np.zeros(4, np.int32), 0, 10, 1, 0.05, np.zeros(2, np.uint32))
    assert len(a) == 0


def test_table_capacity_error_is_loud():
    A = _arks()
    rng = np.random.default_rng(1)
    seq = synth.ACGT[rng.integers(0, 4, 20000)]
    idx = A.ArksIndex(31, 16)  # room for ~1k keys only
    idx.add_ends(seq, np.array([0, len(seq)], dtype=np.uint64), np.array([1], dtype=np.uint32))
    with pytest.raises(A.ArksError) as e:
        idx.finalize()
    assert e.value.code == -4


def test_conreci_remap_and_imap_add():
    """contigs that share a FASTA name are tallied under the first one; arks_imap_add merges external rows"""
    A = _arks()
    k, j = 30, 0.4
    rng = np.random.default_rng(8)
    genome, contigs = synth.make_draft(rng, 100000, 10000, k, n_runs=0, palindromes=0, iupac=0)
    bases, end_off, conreci, names = synth.contig_end_arrays(genome, contigs, k, end_length=4000)
    names = list(names)
    names[3] = names[1]  # duplicate name
    ids, uniq = glue.name_ids(names)
    remap = np.zeros(2 * len(names) + 1, dtype=np.uint32)
    first = {n: names.index(n) for n in names}
    for i, n in enumerate(names):
        remap[2 * i + 1] = 2 * first[n] + 1
        remap[2 * i + 2] = 2 * first[n] + 2
    idx = A.ArksIndex(k, int(end_off[-1]))
    idx.add_ends(bases, end_off, conreci)
    idx.finalize()
    idx.set_conreci_remap(remap)
    rb, roff, bc = synth.make_reads(rng, genome, n_barcodes=40, pairs_per_barcode=50, mol_len=25000, mols_per_barcode=2)
    got = idx.map_pairs(rb, roff, bc, j)
    want_rows = {}
    for b_, c_ in zip(bc.tolist(), got.tolist()):
        if c_:
            r = int(remap[c_])
            ht = want_rows.setdefault((b_, (r - 1) // 2), [0, 0])
            ht[0 if r & 1 else 1] += 1
    gb, gc, gh, gt = idx.imap()
    assert {(b_, c_): [h, t] for b_, c_, h, t in zip(gb.tolist(), gc.tolist(), gh.tolist(), gt.tolist())} == want_rows
    assert not any(c_ == 3 for c_ in gc.tolist())  # the duplicate never appears under its own index
    idx.imap_add([0, 1000], [0, 2], [5, 0], [0, 7])
    gb, gc, gh, gt = idx.imap()
    rows = {(b_, c_): [h, t] for b_, c_, h, t in zip(gb.tolist(), gc.tolist(), gh.tolist(), gt.tolist())}
    assert rows[(1000, 2)] == [0, 7]
    assert rows[(0, 0)] == [want_rows.get((0, 0), [0, 0])[0] + 5, want_rows.get((0, 0), [0, 0])[1]]
