#!/usr/bin/env python
"""Generates tests/golden/cli_cases/: small adversarial FASTA/FASTQ inputs plus the outputs of the
reference's own hot-path code (oracle/_ref/arcs_ref) on them.  Run here (needs /root/reference to have
built oracle/_ref); the fixtures are committed so the GPU box does not need the reference.

  python tools/make_fixtures.py
"""
import gzip
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "cli_cases")
REF = os.path.join(ROOT, "oracle", "_ref", "arcs_ref")


def wrap(seq, width):
    return "\n".join(seq[i:i + width] for i in range(0, len(seq), width))


def make_case(name, seed, k, contig_names=None, fastq_mutator=None, genome_len=90000, mean_contig=9000,
              n_barcodes=50, ppb=40, read_len=150, jitter=30, mol_len=25000, mols=2):
    rng = np.random.default_rng(seed)
    genome, contigs = synth.make_draft(rng, genome_len, mean_contig, k, n_runs=8, palindromes=4, iupac=4)
    d = os.path.join(OUT, name)
    os.makedirs(d, exist_ok=True)
    # FASTA: multi-line, some CRLF, some lower case, comments, a contig shorter than -z, duplicate names
    with open(os.path.join(d, "draft.fa"), "w", newline="") as f:
        for i, (cname, s, e) in enumerate(contigs):
            nm = contig_names[i] if contig_names and i < len(contig_names) else cname
            seq = genome[s:e].tobytes().decode()
            eol = "\r\n" if i % 4 == 1 else "\n"
            f.write(">%s some comment%s" % (nm, eol))
            f.write(wrap(seq, 70 if i % 2 else 10 ** 9).replace("\n", eol) + eol)
            if i % 3 == 0:
                f.write(eol)
        f.write(">tiny\nACGTACGTACGTACGTACGT\n")
    rb, roff, bc = synth.make_reads(rng, genome, n_barcodes=n_barcodes, pairs_per_barcode=ppb, read_len=read_len,
                                    mol_len=mol_len, mols_per_barcode=mols, sub_rate=0.003, n_rate=0.002, len_jitter=jitter)
    recs = []
    for p in range(len(bc)):
        code = "".join("ACGT"[(int(bc[p]) >> (2 * (7 - q))) & 3] for q in range(8)) + "-1"
        for m in range(2):
            s = rb[roff[2 * p + m]:roff[2 * p + m + 1]].tobytes().decode()
            recs.append(["r%d/%d" % (p, m + 1), "BX:Z:%s" % code, s, "I" * len(s)])
    if fastq_mutator:
        recs = fastq_mutator(recs, rng)
    with gzip.open(os.path.join(d, "reads.fq.gz"), "wt", newline="") as f:
        for nm, cm, s, q in recs:
            f.write("@%s%s\n%s\n+\n%s\n" % (nm, (" " + cm) if cm is not None else "", s, q))
    return d


def mutate(recs, rng):
    n = len(recs)
    out = []
    for i in range(0, n, 2):
        a, b = list(recs[i]), list(recs[i + 1])
        t = (i // 2) % 29
        if t == 1:
            a[1] = None  # no comment at all
        elif t == 2:
            b[1] = "RG:Z:x"  # no BX on mate 2
        elif t == 3:
            b[1] = b[1].replace("-1", "-2")  # barcodes differ
        elif t == 4:
            a[0], b[0] = "q%d" % i, "w%d" % i  # unpaired names
        elif t == 5:
            a[1] = "XY:Z:1 " + a[1] + " ZZ:i:3"  # BX in the middle, space terminated
            b[1] = "XY:Z:1 " + b[1] + " ZZ:i:3"
        elif t == 6:
            a[1] = a[1] + "\tQT:Z:x"  # a tab does not terminate the barcode
        elif t == 7:
            a[0], b[0] = "name%d" % i, "name%d" % i  # no /1 /2 suffix
        elif t == 8:
            a[2] = a[2][:10]  # shorter than k
            a[3] = a[3][:10]
        elif t == 9:
            a[2] = a[2][:40] + "NNNNNNNN" + a[2][48:]  # too many N: invalid pair
        elif t == 10:
            a[2] = a[2][:30] + "R" + a[2][31:]  # IUPAC in a read: invalid pair
        elif t == 11:
            a[2] = a[2].lower()
        elif t == 12:
            a[1] = "BX:Z:"  # empty barcode
            b[1] = "BX:Z:"
        elif t == 13:
            a[0] = a[0].replace("/1", "/1x")  # "/" followed by a digit: suffix stripped from there
        out += [a, b]
    # an odd trailing record is dropped by the reference
    out.append(["tail/1", "BX:Z:AAAAAAAA-1", "ACGTACGTACGTACGTACGTACGTACGTACGT", "I" * 32])
    return out


def run_ref(d, args, tag, multfile=None):
    base = os.path.join(d, "expected_" + tag)
    cmd = [REF, "-f", os.path.join(d, "draft.fa"), "-b", base, "--tsv", base + "_main.tsv", "--barcode-counts", base + "_bc.tsv",
           "--dump-imap", base + "_imap.txt", "--dump-pmap", base + "_pmap.txt", "--timing-json", base + "_stats.json"] + args
    if multfile:
        cmd += ["-u", multfile]
    if "-D" in args:  # distance estimation: estimates per edge + the intra-contig samples
        cmd += ["--dist_tsv", base + "_dist.tsv", "--samples_tsv", base + "_samples.tsv"]
    cmd.append(os.path.join(d, "reads.fq.gz"))
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    json.dump({"args": args, "multfile": os.path.basename(multfile) if multfile else None}, open(base + "_args.json", "w"))


def shuffle_pairs(recs, rng):
    """stLFR-style input (BASELINE.json configs[2]): read pairs of all barcodes interleaved at random"""
    order = rng.permutation(len(recs) // 2)
    out = []
    for i in order:
        out += [recs[2 * i], recs[2 * i + 1]]
    return out


def add_shuffled_case():
    # configs[2] shape: k=40, barcodes not grouped in the file, multiplicity filter -m 2-10000
    d = make_case("shuffled_k40", 14, 40, fastq_mutator=shuffle_pairs, genome_len=100000, mean_contig=12000, n_barcodes=400,
                  ppb=12, mol_len=14000, mols=1)
    run_ref(d, ["-k", "40", "-j", "0.5", "-c", "2", "-m", "2-10000", "-e", "30000", "-z", "500", "-r", "0.05", "-t", "1"], "a")
    run_ref(d, ["-k", "40", "-j", "0.5", "-c", "3", "-m", "30-10000", "-e", "2000", "-z", "500", "-r", "0.05", "-t", "1", "-D"], "d")


def add_dist_cases():
    """-D runs of the reference's code on the existing inputs (tags d*, added after the first fixtures)"""
    d = os.path.join(OUT, "mixed_k30")
    run_ref(d, ["-k", "30", "-j", "0.5", "-c", "3", "-m", "20-10000", "-e", "3000", "-z", "500", "-r", "0.05", "-t", "1", "-D", "-B", "3"],
            "d")
    run_ref(d, ["-k", "30", "-j", "0.2", "-c", "2", "-m", "1-1000", "-e", "1500", "-z", "3000", "-r", "0.2", "-l", "2", "-d", "3", "-t", "1",
                "-D"], "e")
    d = os.path.join(OUT, "plain_k60")
    run_ref(d, ["-k", "60", "-j", "0.55", "-c", "5", "-m", "50-10000", "-e", "4000", "-z", "500", "-r", "0.05", "-t", "1", "-D", "-B", "2"],
            "d")
    d = os.path.join(OUT, "long_k20")
    run_ref(d, ["-k", "20", "-j", "0.05", "-c", "3", "-m", "8-10000", "-e", "5000", "-z", "500", "-r", "0.05", "-t", "1", "-D"], "d")


def make_sam_case():
    """ARCS alignment mode (Arcs.cpp:572-771): synthetic SAM text with consecutive read pairs and the record
    shapes the state machine distinguishes; expected outputs from the reference's own readBAM code"""
    rng = np.random.default_rng(21)
    d = os.path.join(OUT, "sam_arcs")
    os.makedirs(d, exist_ok=True)
    lens = [9000, 14000, 400, 7000, 30000, 12000, 8000, 5200]
    names = ["10", "9", "tiny", "2", "b", "a", "100", "1"]
    acgt = "ACGT"
    with open(os.path.join(d, "draft.fa"), "w") as f:
        for nm, L in zip(names, lens):
            f.write(">%s len=%d\n%s\n" % (nm, L, "".join(acgt[int(x)] for x in rng.integers(0, 4, L))))
    lines = ["@HD\tVN:1.5\tSO:unsorted"] + ["@SQ\tSN:%s\tLN:%d" % (nm, L) for nm, L in zip(names, lens)] + ["@PG\tID:bwa\tPN:bwa"]
    n_bc, ppb = 90, 30
    t = 0
    for b in range(n_bc):
        code = "".join(acgt[(b >> (2 * (7 - q))) & 3] for q in range(8))
        c1 = int(rng.integers(0, len(names)))
        c2 = (c1 + 1 + int(rng.integers(0, 2))) % len(names)  # a molecule spans the ends of two contigs
        ends = {c1: rng.random() < 0.5, c2: rng.random() < 0.5}
        for p in range(ppb):
            c = c1 if p % 2 == 0 else c2
            L = lens[c]
            end_head = ends[c] if rng.random() < 0.9 else not ends[c]
            pos = int(rng.integers(1, max(2, min(2500, L // 2)))) if end_head else int(L - rng.integers(min(150, L // 4), max(min(150, L // 4) + 1, min(2600, L // 2))))
            pos2 = pos + int(rng.integers(150, 400))
            fl = (99, 147) if rng.random() < 0.5 else (83, 163)
            rn = "r%d" % t
            tags = "NM:i:%d\tBX:Z:%s-1\tQT:Z:IIII" % (int(rng.integers(0, 3)), code)
            cig = ["150M", "100M2I48M", "5S145M", "60M1D90M", "75=1X74="][int(rng.integers(0, 5))]
            seq = "".join(acgt[int(x)] for x in rng.integers(0, 4, 150))
            rec = [[rn, fl[0], names[c], pos, 60, cig, "=", pos2, 300, seq, "I" * 150, tags],
                   [rn, fl[1], names[c], pos2, 60, "150M", "=", pos, -300, seq[::-1], "I" * 150, tags]]
            k = t % 23
            if k == 1:
                for r in rec:
                    r[0], r[11] = "q%d_%s" % (t, code), "NM:i:1"  # barcode from the read name
            elif k == 2:
                for r in rec:
                    r[0], r[11] = "q%d_ACGTNX" % t, "NM:i:1"  # not a barcode
            elif k == 3:
                rec[0][4] = 0  # MAPQ 0
            elif k == 4:
                rec[1][11] = "NM:i:25\tBX:Z:%s-1" % code  # low identity
            elif k == 5:
                rec[1][2] = names[(c + 1) % len(names)]  # mates on different contigs
            elif k == 6:
                rec[0][2] = rec[1][2] = "*"
            elif k == 7:
                rec.append([rn, 2048 + fl[0], names[c], pos + 50, 60, "80M70H", "=", pos2, 0, seq[:80], "I" * 80, tags])  # third line
            elif k == 8:
                rec = rec[:1]  # singleton
            elif k == 9:
                rec[0][1], rec[1][1] = 65, 129  # not proper pairs
            elif k == 10:
                rec[0][1] |= 256  # secondary: not counted in the multiplicity, pair rejected by the flag test
            elif k == 11:
                rec[0][3], rec[1][3] = L // 2 - 100, L // 2 + 100  # middle of the contig
            elif k == 12:
                rec[1][11] = "XA:Z:foo BX:Z:%s-1 XT:i:0" % code  # blank-separated tags
            elif k == 13:
                rec[0][5] = "*"  # no CIGAR: identity 0
            for r in rec:
                lines.append("\t".join(str(x) for x in r))
            t += 1
    lines.append("")
    with open(os.path.join(d, "aln.sam"), "w") as f:
        f.write("\n".join(lines))
    for tag, args, with_f in (
            ("a", ["-s", "98", "-c", "2", "-m", "4-10000", "-e", "3000", "-z", "500", "-r", "0.05", "-l", "0"], True),
            ("b", ["-s", "90", "-c", "3", "-m", "10-10000", "-e", "0", "-z", "1000", "-r", "0.1", "-l", "2", "-d", "3", "-D", "-B", "4"], False)):
        base = os.path.join(d, "expected_" + tag)
        cmd = [REF, "--arcs", "-b", base, "--tsv", base + "_main.tsv", "--barcode-counts", base + "_bc.tsv", "--dump-imap",
               base + "_imap.txt", "--dump-pmap", base + "_pmap.txt", "--timing-json", base + "_stats.json"] + args
        if with_f:
            cmd += ["-f", os.path.join(d, "draft.fa")]
        if "-D" in args:
            cmd += ["--dist_tsv", base + "_dist.tsv", "--samples_tsv", base + "_samples.tsv"]
        cmd.append(os.path.join(d, "aln.sam"))
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        json.dump({"args": args, "multfile": None, "mode": "arcs", "with_f": with_f}, open(base + "_args.json", "w"))


def make_cut_case():
    """arks-long without the pipe: long reads in, `arcs --arks --cut 250` must give what the
    reference's code gives on the output of long-to-linked-pe (our drop-in of that tool is pinned byte for byte
    on the reference's own golden, tests/test_long_to_linked_pe.py)."""
    import tempfile
    ltlpe = os.path.join(ROOT, "arcs_b200", "bin", "long-to-linked-pe")
    rng = np.random.default_rng(21)
    k = 20
    genome, contigs = synth.make_draft(rng, 120000, 10000, k, n_runs=6, palindromes=3, iupac=3)
    d = os.path.join(OUT, "cut_k20")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "draft.fa"), "w") as f:
        for cname, s, e in contigs:
            f.write(">%s\n%s\n" % (cname, wrap(genome[s:e].tobytes().decode(), 100)))
    comp = bytes.maketrans(b"ACGTacgtNn", b"TGCAtgcaNn")
    with gzip.open(os.path.join(d, "long_reads.fa.gz"), "wt") as f:
        for i in range(130):
            L = int(np.clip(rng.lognormal(np.log(6000), 0.7), 400, 30000))  # some are shorter than --cut_min
            a = int(rng.integers(0, len(genome) - L))
            r = genome[a:a + L].copy()
            sub = rng.random(L) < 0.03  # long-read error rate (substitutions only: the cut positions stay put)
            r[sub] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(sub.sum()))]
            if i % 9 == 0:  # N-rich stretch: some pseudo reads exceed the 2 % N limit
                p = int(rng.integers(0, L - 300))
                r[p:p + 300][rng.random(300) < 0.1] = ord("N")
            if i % 11 == 3:
                r[int(rng.integers(0, L))] = ord("R")  # IUPAC: that pseudo pair is invalid
            seq = r.tobytes()
            if i % 2:
                seq = seq.translate(comp)[::-1]
            seq = seq.decode()
            if i % 5 == 4:  # a FASTQ record among the FASTA ones
                f.write("@lr%d extra words\n%s\n+\n%s\n" % (i, seq, "".join(chr(33 + (j * 7 + i) % 40) for j in range(L))))
            else:
                f.write(">lr%d len=%d\n%s\n" % (i, L, wrap(seq, 80)))
    reads = os.path.join(d, "long_reads.fa.gz")
    with tempfile.TemporaryDirectory() as t:
        cut = os.path.join(t, "cut.fq")
        with open(cut, "wb") as o:
            subprocess.check_call([ltlpe, "-l", "250", "-m", "2000", reads], stdout=o, stderr=subprocess.DEVNULL)
        mult = os.path.join(d, "mult.tsv")
        subprocess.check_call([ltlpe, "-l", "250", "-m", "2000", "--bx-only", "-b", mult, reads], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
        # drop a few barcodes from the multiplicity file: their pairs are rejected as invalid barcodes
        lines = open(mult).read().splitlines()
        open(mult, "w").write("\n".join(ln for i, ln in enumerate(lines) if i % 13 != 5) + "\n")
        for tag, args, mf in (("a", ["-k", "20", "-j", "0.05", "-c", "2", "-m", "4-10000", "-e", "30000", "-z", "500", "-r", "0.05", "-t", "1"], None),
                              ("u", ["-k", "20", "-j", "0.05", "-c", "3", "-m", "8-10000", "-e", "3000", "-z", "500", "-r", "0.05", "-t", "1"], mult)):
            base = os.path.join(d, "expected_" + tag)
            cmd = [REF, "-f", os.path.join(d, "draft.fa"), "-b", base, "--tsv", base + "_main.tsv", "--barcode-counts", base + "_bc.tsv",
                   "--dump-pmap", base + "_pmap.txt"] + args + (["-u", mf] if mf else []) + [cut]
            subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            for junk in (base + ".dist.gv",):
                if os.path.exists(junk):
                    os.remove(junk)
            json.dump({"args": args, "multfile": "mult.tsv" if mf else None, "cut": [250, 2000], "reads": "long_reads.fa.gz"},
                      open(base + "_args.json", "w"))
    return d


def main():
    if not os.path.exists(REF):
        raise SystemExit("oracle/_ref/arcs_ref is missing: run oracle/build_ref.sh where /root/reference exists")
    if len(sys.argv) > 1 and sys.argv[1] == "dist":
        return add_dist_cases()
    if len(sys.argv) > 1 and sys.argv[1] == "sam":
        return make_sam_case()
    if len(sys.argv) > 1 and sys.argv[1] == "cut":
        return make_cut_case()
    if len(sys.argv) > 1 and sys.argv[1] == "shuffled":
        return add_shuffled_case()
    # case A: k=30 defaults-ish, adversarial FASTQ, contig names whose string order differs from numeric order
    names = ["10", "9", "100", "2", "b", "a", "2", "11", "1", "3"]  # "2" appears twice
    d = make_case("mixed_k30", 11, 30, contig_names=names, fastq_mutator=mutate, n_barcodes=120, ppb=60, mol_len=9000, mols=1)
    run_ref(d, ["-k", "30", "-j", "0.5", "-c", "3", "-m", "20-10000", "-e", "3000", "-z", "500", "-r", "0.05", "-t", "1"], "a")
    run_ref(d, ["-k", "30", "-j", "0.2", "-c", "2", "-m", "1-1000", "-e", "0", "-z", "3000", "-r", "0.2", "-l", "2", "-d", "2", "-t", "1"], "b")
    # with a multiplicity file that lacks some barcodes (they are rejected as invalid)
    with open(os.path.join(d, "mult.csv"), "w") as f:
        bcs = sorted({l.split("\t")[0] for l in open(os.path.join(d, "expected_a_bc.tsv"))})
        for i, b in enumerate(bcs):
            if i % 5:
                f.write("%s,%d\n" % (b, 40 + i))
    run_ref(d, ["-k", "30", "-j", "0.5", "-c", "2", "-m", "45-200", "-e", "3000", "-z", "500", "-r", "0.05", "-t", "1"], "c",
            multfile=os.path.join(d, "mult.csv"))
    # case B: k=60 (two-word keys), longer ends
    d = make_case("plain_k60", 12, 60, genome_len=120000, mean_contig=15000, n_barcodes=150, ppb=60, mol_len=12000, mols=1)
    run_ref(d, ["-k", "60", "-j", "0.55", "-c", "5", "-m", "50-10000", "-e", "30000", "-z", "500", "-r", "0.05", "-t", "1"], "a")
    # case C: k=20, low Jaccard threshold, long-read style pseudo pairs (250 bp)
    d = make_case("long_k20", 13, 20, genome_len=100000, mean_contig=20000, n_barcodes=300, ppb=12, read_len=250, jitter=0, mol_len=15000, mols=1)
    run_ref(d, ["-k", "20", "-j", "0.05", "-c", "3", "-m", "8-10000", "-e", "30000", "-z", "500", "-r", "0.05", "-t", "1"], "a")
    add_dist_cases()
    make_sam_case()
    add_shuffled_case()
    print("fixtures written under", OUT)
    subprocess.call(["du", "-sh", OUT])


if __name__ == "__main__":
    main()
