"""Pins the CPU restatement (oracle/arks_oracle.c) against the reference:
 - the reference's unmodified ReadsProcessor::prepSeq, window by window (where oracle/_ref exists),
 - known-answer keys quoted in SURVEY.md section 8(a-1),
 - the golden outputs + verbose counters of Examples/arks_test-demo and arks-long_test-demo."""
import os

import numpy as np
import pytest

import glue
import oracle_lib as O
import seqio

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_key_known_answers():
    # observed on the reference: palindromes get a deterministic garbage key
    assert O.key(b"TACGTTTTACGTAAAACGTA", 20).hex() == "c6ff1b0000"
    w30 = b"TACGTTTTACGTCGAACGTAAAACGTA"
    assert O.key(b"ACGT", 4) is not None
    assert O.key(b"ACGN", 4) is None
    assert O.key(b"acgtacgtaa", 10) == O.key(b"ACGTACGTAA", 10)
    # canonical = min(fwd, revcomp)
    assert O.key(b"TTTTTTTT", 8) == bytes([0, 0])
    assert O.key(b"AAAAAAAC", 8) == bytes([0x00, 0x01])
    assert O.key(b"GTTTTTTT", 8) == bytes([0x00, 0x01])
    del w30


def _rand_seq(rng, n, p_n=0.01, p_lower=0.1):
    s = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
    lower = rng.random(n) < p_lower
    s = np.where(lower, s | 0x20, s)
    bad = rng.random(n) < p_n
    s = np.where(bad, rng.choice(np.frombuffer(b"NnRY-*x", dtype=np.uint8), size=n), s)
    return s.astype(np.uint8).tobytes()


def _revcomp(s):
    return s.translate(bytes.maketrans(b"ACGTacgt", b"TGCAtgca"))[::-1]


@pytest.mark.skipif(O.ref_prepseq_lib() is None, reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("k", [4, 5, 7, 8, 9, 12, 14, 15, 16, 17, 20, 21, 24, 28, 30, 31, 32, 33, 40, 59, 60, 63, 64, 65, 96, 100, 128])
def test_key_matches_reference_prepseq(k):
    rng = np.random.default_rng(k)
    seqs = [_rand_seq(rng, 3000)]
    # palindromes (even k) and near-palindromes, (AT)n / (CG)n tracts
    if k % 2 == 0 and k not in (6, 10):
        for _ in range(200):
            half = _rand_seq(rng, k // 2, p_n=0, p_lower=0.2)
            seqs.append(b"GA" + half + _revcomp(half) + b"TC")
        seqs.append(b"AT" * (k + 5))
        seqs.append(b"CG" * (k + 5))
        seqs.append(b"ACGT" * (k + 5))
    for s in seqs:
        keys, valid = O.ref_keys_all(s, k)
        for i in range(len(valid)):
            mine = O.key(s[i:i + k], k)
            if valid[i]:
                assert mine is not None and mine == keys[i].tobytes(), (k, i, s[i:i + k])
            else:
                assert mine is None, (k, i, s[i:i + k])


def _run_demo(fa, fq, k, j, c, m, e, z, r, l, multfile=None):
    contigs = seqio.read_fasta(fa)
    ends, names = glue.contig_ends(contigs, z, e)
    km = O.KMap(k, sum(len(s) for s, _ in ends))
    for s, conreci in ends:
        km.map_kmers(s, conreci)
    recs = seqio.read_fastq(fq)
    if multfile:
        mult = {}
        for line in open(multfile):
            b, n = line.split()
            mult[b] = int(n)
    else:
        mult = seqio.multiplicities(recs)
    barcodes, bases, off = seqio.candidate_pairs(recs, mult)
    conreci, st = km.map_pairs(bases, off, j)
    imap, uniq = glue.imap_rows(barcodes, conreci, names)
    rank = glue.lex_rank(uniq)
    bc, ct, hd, tl, mu, _ = glue.imap_to_arrays(imap, mult)
    a, b, counts = O.pair_contigs(bc, ct, hd, tl, mu, m[0], m[1], c, r, rank)
    gv = glue.gv_text(a, b, counts, uniq, rank, l, r)
    return km, st, gv, (a, b, counts), imap


def test_arks_demo_golden():
    d = os.path.join(GOLD, "arks_demo")
    km, st, gv, _, imap = _run_demo(os.path.join(d, "test_scaffolds.renamed.fa"), os.path.join(d, "test_reads.fq.gz"),
                                   k=30, j=0.55, c=5, m=(50, 6000), e=30000, z=500, r=np.float32(0.05), l=0)
    assert gv == open(os.path.join(d, "expected_original.gv")).read()
    # verbose counters of the golden log (expected_arks_log_excerpt.txt)
    s = km.stats.as_dict()
    assert (s["kmers_valid"], s["kmers_null"], s["recorded"], s["collisions"], s["removed"], s["unique"]) == (
        123190, 303, 118710, 4480, 547, 118334)
    t = st.as_dict()
    assert (t["pairs_stored"], t["pairs_invalid"], t["pairs_nogood"]) == (21632, 0, 6012)
    assert (t["kmers_valid"], t["kmers_invalid"], t["found"], t["recorded"], t["dups"]) == (
        6109324, 0, 4862376, 4814099, 48277)
    assert (t["reads_pass"], t["reads_fail"]) == (44503, 10785)
    assert len(imap) == 625


def test_arks_long_demo_golden():
    d = os.path.join(GOLD, "arks_long_demo")
    _, _, gv, _, _ = _run_demo(os.path.join(d, "test_scaffolds.renamed.fa"), os.path.join(d, "test_reads.cut250.fq.gz"),
                              k=20, j=0.05, c=3, m=(8, 10000), e=30000, z=500, r=np.float32(0.05), l=0,
                              multfile=os.path.join(d, "barcodeMultiplicityArcs.tsv"))
    assert gv == open(os.path.join(d, "expected_original.gv")).read()


def test_head_or_tail_examples():
    # r=0.05 -> minMax[sum]: 3->3, 5->5, 6->6, 7->6, 10->8, 20->14, 30->20
    r = np.float32(0.05)
    for s, mm in [(3, 3), (5, 5), (6, 6), (7, 6), (10, 8), (20, 14), (30, 20)]:
        assert O.head_or_tail(mm, s - mm, 1, r)[0]
        assert not O.head_or_tail(mm - 1, s - mm + 1, 1, r)[0] or mm - 1 < s - mm + 1
    assert O.head_or_tail(5, 5, 1, r) == (False, False)
    assert O.head_or_tail(9, 0, 5, r) == (True, True)
    assert O.head_or_tail(0, 9, 5, r) == (True, False)
    assert O.head_or_tail(4, 0, 5, r) == (False, False)  # below min_reads
