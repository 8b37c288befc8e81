"""long-to-linked-pe (arcs_b200/host/long_to_linked_pe.cpp) against the reference's own golden:
Examples/arks-long_test-demo/test_reads.fa.gz -> output/test_reads.cut250.fq.gz.

The golden was produced before the tool grew its -m filter (it contains pseudo reads of long reads
shorter than 2000 bp, e.g. record 13 with 1851 bp), so it is reproduced with `-l 250 -m 0`; the -m
rule itself (src/long-to-linked-pe.cpp:204,223) is checked on the same data by construction.
CPU only: the tool is host code."""
import gzip
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "arcs_b200", "bin", "long-to-linked-pe")
GOLD = os.path.join(ROOT, "tests", "golden", "arks_long_demo")
REF_DEMO = "/root/reference/Examples/arks-long_test-demo"

pytestmark = pytest.mark.skipif(not os.path.exists(TOOL), reason="host tools not built (run __graft_entry__.build())")


def _run(args, cwd):
    return subprocess.run([TOOL] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)


def _golden_prefix(n_records):
    """the part of the golden that belongs to the first n_records long reads (barcodes 1..n)"""
    data = gzip.open(os.path.join(GOLD, "test_reads.cut250.fq.gz"), "rb").read()
    m = re.search(rb"^@\S+ BX:Z:%d\n" % (n_records + 1), data, re.M)
    return data[:m.start()] if m else data


def _records(fa_gz):
    recs, name, seq = [], None, []
    for line in gzip.open(fa_gz, "rt"):
        line = line.rstrip("\n")
        if line.startswith(">"):
            if name is not None:
                recs.append((name, "".join(seq)))
            name, seq = line[1:].split()[0], []
        else:
            seq.append(line)
    recs.append((name, "".join(seq)))
    return recs


def test_first_60_long_reads_match_the_golden(tmp_path):
    out = _run(["-l", "250", "-m", "0", os.path.join(GOLD, "test_reads.head60.fa.gz")], tmp_path).stdout
    assert out == _golden_prefix(60)


@pytest.mark.skipif(not os.path.exists(REF_DEMO), reason="reference tree not present")
def test_whole_demo_matches_the_golden(tmp_path):
    out = _run(["-l", "250", "-m", "0", "-t", "8", os.path.join(REF_DEMO, "test_reads.fa.gz")], tmp_path).stdout
    assert out == gzip.open(os.path.join(REF_DEMO, "output", "test_reads.cut250.fq.gz"), "rb").read()


def test_min_length_filter_and_multiplicities(tmp_path):
    fa = os.path.join(GOLD, "test_reads.head60.fa.gz")
    recs = _records(fa)
    l, m = 250, 2000
    out = _run(["-l", str(l), "-m", str(m), "--bx", "-b", "mult.tsv", fa], tmp_path).stdout
    full = _run(["-l", str(l), "-m", "0", fa], tmp_path).stdout
    # -m only removes the barcodes of reads shorter than m
    keep = {i + 1 for i, (_, s) in enumerate(recs) if len(s) >= m and len(s) >= 2 * l}
    want = b"".join(rec for rec in re.findall(rb"@\S+ BX:Z:\d+\n[^\n]*\n\+\n[^\n]*\n", full)
                    if int(re.match(rb"@\S+ BX:Z:(\d+)", rec).group(1)) in keep)
    assert out == want
    # multiplicity = number of pseudo reads of the barcode (src/long-to-linked-pe.cpp:207-212)
    rows = [tuple(map(int, ln.split("\t"))) for ln in open(tmp_path / "mult.tsv")]
    counts = {}
    for b in re.findall(rb"BX:Z:(\d+)\n", out):
        counts[int(b)] = counts.get(int(b), 0) + 1
    assert rows == sorted(counts.items())
    only = _run(["-l", str(l), "-m", str(m), "--bx-only", "-b", "mult2.tsv", fa], tmp_path)
    assert only.stdout == b"" and open(tmp_path / "mult2.tsv").read() == open(tmp_path / "mult.tsv").read()


def test_fastq_input_fasta_output_and_revcomp_table(tmp_path):
    seq = "ACGTNRYKMSWBDHVacgtnu" * 30  # 630 bases, IUPAC + lower case
    qual = "".join(chr(33 + (i % 40)) for i in range(len(seq)))
    (tmp_path / "in.fq").write_text("@r1 some comment\n%s\n+\n%s\n@short\nACGT\n+\nIIII\n" % (seq, qual))
    comp = dict(zip("ACGTURYSWKMBDHVNacgturyswkmbdhvn", "TGCAAYRSWMKVHDBNtgcaayrswmkvhdbn"))

    def rc(s):
        return "".join(comp[c] for c in reversed(s))

    out = _run(["-l", "100", "-m", "0", "in.fq"], tmp_path).stdout.decode()
    exp = []
    for n, i in enumerate(range(0, len(seq) - 200 + 1, 200), 1):
        exp += ["@r1_f%d BX:Z:1" % n, seq[i:i + 100], "+", qual[i:i + 100],
                "@r1_f%d BX:Z:1" % n, rc(seq[i + 100:i + 200]), "+", qual[i + 100:i + 200][::-1]]
    rem = len(seq) % 200  # 30: forward piece = the remainder, reverse piece = the same bases
    exp += ["@r1_f4 BX:Z:1", seq[-rem:], "+", qual[-rem:], "@r1_f4 BX:Z:1", rc(seq[-rem:]), "+", qual[-rem:][::-1]]
    assert out == "\n".join(exp) + "\n"
    fa = _run(["-l", "100", "-m", "0", "--fasta", "in.fq"], tmp_path).stdout.decode().splitlines()
    assert fa == [(">" + x[1:]) if j % 4 == 0 else x for j, x in enumerate(exp) if j % 4 < 2]


def test_usage_errors(tmp_path):
    p = subprocess.run([TOOL, "-m", "5", "x.fa"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode != 0 and b"missing option -- 'l'" in p.stderr
    p = subprocess.run([TOOL, "-l", "250"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode != 0 and b"missing file operand" in p.stderr


# ---- `arcs --arks --cut L`: the same bytes without the tool and the pipe (arcs_b200/host/long_cut.h) ---------
INGEST = os.path.join(ROOT, "arcs_b200", "bin", "ingest_dump")


@pytest.mark.parametrize("reads,l,m", [
    (os.path.join(ROOT, "tests", "golden", "cli_cases", "cut_k20", "long_reads.fa.gz"), 250, 2000),
    (os.path.join(ROOT, "tests", "golden", "cli_cases", "cut_k20", "long_reads.fa.gz"), 100, 0),
    (os.path.join(GOLD, "test_reads.head60.fa.gz"), 250, 2000),
    (os.path.join(GOLD, "test_reads.head60.fa.gz"), 1500, 500),
])
def test_cutting_inside_the_ingest_equals_the_pipe(reads, l, m, tmp_path):
    """what the read ingest of the CLI sees with --cut (pairs, barcodes, counters, multiplicities) is what it
    sees when the tool's output is fed to it, for the sequential reader and the block-parallel one"""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "arcs_b200", "host"), "../bin/ingest_dump", "../bin/long-to-linked-pe"])
    cut = tmp_path / "cut.fq"
    with open(cut, "wb") as o:
        subprocess.check_call([TOOL, "-l", str(l), "-m", str(m), reads], stdout=o, stderr=subprocess.DEVNULL)
    want = subprocess.run([INGEST, "seq", str(cut)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    assert want.count(b"\n") > 50
    env = dict(os.environ, ARKS_CUT="%d,%d" % (l, m))
    got = subprocess.run([INGEST, "seq", reads], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True, env=env).stdout
    assert got == want
    for workers, block in ((1, 1 << 16), (4, 1 << 15), (3, 1 << 20)):
        p = subprocess.run([INGEST, "par", reads, str(workers), str(block)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True, env=env)
        assert p.stdout == want, (workers, block)
        if block < (1 << 20):  # the cut reads are strict four-line FASTQ: the block parser takes them
            assert int(p.stderr.decode().split("FAST_BLOCKS")[1].split()[0]) > 0


def test_cut_option_checks_of_the_cli(tmp_path):
    """--cut is an --arks option like -k / -j / -t (Arcs.cpp:2087-2091), and only with it are FASTA names
    accepted as read files (checkSameFormat, Arcs.cpp:336-361, knows .fq / .fastq only)"""
    arcs = os.path.join(ROOT, "arcs_b200", "bin", "arcs")
    d = os.path.join(ROOT, "tests", "golden", "cli_cases", "cut_k20")
    run = lambda args: subprocess.run([arcs] + args, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    p = run(["--cut", "250", "-f", os.path.join(d, "draft.fa"), os.path.join(d, "long_reads.fa.gz")])
    assert p.returncode != 0 and "does not match with method" in p.stderr
    p = run(["--arks", "-k", "20", "-f", os.path.join(d, "draft.fa"), "-b", "x", os.path.join(d, "long_reads.fa.gz")])
    assert p.returncode != 0 and "Unknown type file" in p.stdout
    p = run(["--arks", "--cut", "250", "-k", "20", "-f", os.path.join(d, "draft.fa"), "-b", "x", os.path.join(d, "long_reads.fa.gz")])
    # past the option checks: either it runs (GPU present) or the GPU initialisation fails loudly
    assert "Unknown type file" not in p.stdout and "does not match" not in p.stderr
    assert p.returncode == 0 or "no CPU fallback" in p.stderr


def test_span_and_dist_parameters(tmp_path):
    """-s / -d append `span` and `read_p<P>` to the parameter file (src/long-to-linked-pe.cpp:294-322): span =
    (total bases / g, integer division) * c; dist = the P-th percentile of the read lengths above 1000, the mean
    of two neighbours when the rank is whole.  `-t N` also names the parameter file (its case has no break)."""
    lens = [400, 1200, 1500, 2600, 3000, 5200, 900, 7000]
    with open(tmp_path / "in.fa", "w") as f:
        for i, n in enumerate(lens):
            f.write(">r%d\n%s\n" % (i, "ACGT" * (n // 4)))
    total = sum(lens)
    over = sorted(n for n in lens if n > 1000)  # 6 values
    _run(["-l", "250", "-s", "-g", "1e3", "-c", "0.5", "-d", "-p", "50", "-f", "p.tsv", "in.fa"], tmp_path)
    rank = 0.5 * len(over)  # 3.0, whole: mean of [2] and [3]
    assert open(tmp_path / "p.tsv").read() == "span\t%d\nread_p50\t%d\n" % (int(total // 1000 * 0.5), (over[2] + over[3]) // 2)
    assert rank == int(rank)
    _run(["-l", "250", "-d", "-p", "70", "-f", "p.tsv", "in.fa"], tmp_path)  # appended; rank 4.2 -> [4]
    assert open(tmp_path / "p.tsv").read().splitlines()[2] == "read_p70\t%d" % over[4]
    _run(["-l", "250", "-d", "-t", "3", "in.fa"], tmp_path)  # the parameter file is now called "3"
    assert open(tmp_path / "3").read() == "read_p50\t%d\n" % ((over[2] + over[3]) // 2)
    # with --bx the reads that are too short to be cut do not count for -s / -d (the `continue` of :196)
    _run(["-l", "250", "-s", "-g", "1000", "-d", "--bx", "-b", "m.tsv", "-f", "q.tsv", "in.fa"], tmp_path)
    kept = [n for n in lens if n >= 2000]
    k = sorted(kept)
    r = 0.5 * len(k)
    want = (k[int(r) - 1] + k[int(r)]) // 2 if r == int(r) else k[int(r)]
    assert open(tmp_path / "q.tsv").read() == "span\t%d\nread_p50\t%d\n" % (int(sum(kept) // 1000 * 0.25), want)
    # nothing above 1000 bases: a message instead of an estimate
    (tmp_path / "short.fa").write_text(">a\nACGT\n")
    p = subprocess.run([TOOL, "-l", "2", "-m", "0", "-d", "-f", "e.tsv", "short.fa"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert b"unable to estimate dist parameter" in p.stderr
