"""Host-side logic on CPU: the C++ FASTA/FASTQ reader of the drop-in CLI (arcs_b200/host/seq_reader.h)
against an independent Python restatement of the reference reader's record grammar
(Arcs/kseq.h:175-215), on adversarial text and across buffer sizes."""
import gzip
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DUMP = os.path.join(ROOT, "arcs_b200", "bin", "seq_dump")


def _build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "arcs_b200", "host"), "../bin/seq_dump"])


class KseqModel:
    """the grammar, written from its description"""

    def __init__(self, data: bytes):
        self.d, self.i, self.last = data, 0, 0

    def getc(self):
        if self.i >= len(self.d):
            return -1
        c = self.d[self.i]
        self.i += 1
        return c

    def until(self, line, s=b""):
        """-> (bytes, delimiter or 0, got_any)"""
        if self.i >= len(self.d):
            return s, 0, False
        j = self.i
        if line:
            while j < len(self.d) and self.d[j] != 10:
                j += 1
        else:
            while j < len(self.d) and chr(self.d[j]) not in " \t\n\v\f\r":
                j += 1
        s = s + self.d[self.i:j]
        delim = self.d[j] if j < len(self.d) else 0
        self.i = j + 1
        if line and len(s) > 1 and s.endswith(b"\r"):
            s = s[:-1]
        return s, delim, True

    def read(self):
        if self.last == 0:
            while True:
                c = self.getc()
                if c == -1:
                    return -1, None
                if c in (62, 64):
                    break
            self.last = c
        name, delim, got = self.until(False)
        if not got:
            return -1, None
        comment = b""
        if delim != 10:
            comment, _, _ = self.until(True)
        seq = b""
        while True:
            c = self.getc()
            if c == -1 or c in (62, 43, 64):
                break
            if c == 10:
                continue
            seq, _, _ = self.until(True, seq + bytes([c]))
        if c in (62, 64):
            self.last = c
        if c != 43:
            return len(seq), (name, comment, seq, 0)
        while True:
            c = self.getc()
            if c == -1:
                return -2, None
            if c == 10:
                break
        qual = b""
        while True:
            qual, _, got = self.until(True, qual)
            if not got or len(qual) >= len(seq):
                break
        self.last = 0
        if len(seq) != len(qual):
            return -2, None
        return len(seq), (name, comment, seq, len(qual))


def model_dump(data):
    m = KseqModel(data)
    lines = []
    while True:
        l, rec = m.read()
        if l < 0:
            lines.append("END %d" % l)
            break
        name, comment, seq, ql = rec
        cut = lambda b: b.split(b"\0")[0].decode("latin1")  # noqa: E731
        lines.append("%d\t%s\t%s\t%s\t%d" % (l, cut(name), cut(comment), cut(seq), ql))
    return lines


CASES = {
    "fastq_plain": b"@r1/1 BX:Z:AAA-1\nACGT\n+\nIIII\n@r1/2 BX:Z:AAA-1\nTTGA\n+r1\nIIII\n",
    "fasta_multiline_crlf": b">c1 desc here\r\nACGT\r\nACGT\r\n\r\n>c2\nAC\nGT\n>c3\tx\n\nNNNN",
    "fastq_multiline_at_in_qual": b"@a x\nACGT\nACGT\n+\n@III\nIIII\n@b\nAC\n+\nII\n",
    "junk_before_header": b"junk line\nmore\n@r c\nAC\n+\nII\n",
    "truncated_quality": b"@r\nACGT\n+\nII",
    "no_quality_line": b"@r\nACGT\n+",
    "empty_sequence": b"@e1 BX:Z:X\n\n+\n\n@e2\nAC\n+\nII\n",
    "name_only_at_eof": b">only",
    "tabs_and_cr": b"@n\tBX:Z:Q\tZZ\r\nAC\r\n+\r\nII\r\n",
    "lone_cr_line": b">x\n\r\nAC\n",
    "empty_file": b"",
    "mixed_fastq_then_fasta": b"@q\nACGT\n+\nIIII\n>f\nGGCC\nTT\n",
    "nul_inside": b"@z c\nAC\0GT\n+\nIIIII\n",
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("bufsize", [1, 2, 3, 7, 64, 1 << 20])
@pytest.mark.parametrize("gz", [False, True])
def test_reader_matches_grammar(name, bufsize, gz, tmp_path):
    _build()
    data = CASES[name]
    path = tmp_path / ("in.fq.gz" if gz else "in.fq")
    if gz:
        with gzip.open(path, "wb") as f:
            f.write(data)
    else:
        path.write_bytes(data)
    out = subprocess.run([DUMP, str(path), str(bufsize)], stdout=subprocess.PIPE, check=True).stdout.decode("latin1")
    assert out.split("\n")[:-1] == model_dump(data)


def test_reader_random_text(tmp_path):
    """random mixtures of the grammar's special bytes"""
    _build()
    rng = np.random.default_rng(0)
    alphabet = np.frombuffer(b"@>+\n\n\n\r \tACGTNacgt!I", dtype=np.uint8)
    for t in range(60):
        data = alphabet[rng.integers(0, len(alphabet), int(rng.integers(0, 400)))].tobytes()
        path = tmp_path / ("r%d.txt" % t)
        path.write_bytes(data)
        for bufsize in (5, 1 << 16):
            out = subprocess.run([DUMP, str(path), str(bufsize)], stdout=subprocess.PIPE, check=True).stdout.decode("latin1")
            assert out.split("\n")[:-1] == model_dump(data), data


# ---- the draft on all cores (arcs_b200/host/fasta_fast.h) against the faithful reader ----
def _fast_dump(path, threads):
    return subprocess.run([DUMP, "--fast", str(path), str(threads)], stdout=subprocess.PIPE, check=True).stdout.decode("latin1").split("\n")[:-1]


@pytest.mark.parametrize("threads", [1, 3, 16])
def test_fast_fasta_lists_the_records_the_reader_sees(threads, tmp_path):
    _build()
    rng = np.random.default_rng(threads)
    recs = []
    for i in range(400):
        L = int(rng.choice([0, 1, 2, 5, 59, 60, 61, 600, 5000, 70000]))
        seq = "".join("ACGTNacgtnRY>@+"[int(x)] for x in rng.integers(0, 15, L))
        width = int(rng.choice([0, 60, 61, 7]))
        if width:  # wrapped: a wrapped line must not start with '>', '@' or '+' (that is the non-strict shape, below)
            lines = [seq[a:a + width] for a in range(0, L, width)] or [""]
            lines = [("A" + ln[1:]) if ln[:1] in (">", "@", "+") else ln for ln in lines]
            seq, body = "".join(lines), "\n".join(lines)
        else:
            seq = ("C" + seq[1:]) if seq[:1] in (">", "@", "+") else seq
            body = seq
        name = "c%d" % i if i % 7 else ""
        comment = ["", " len=%d" % L, "\tx y"][i % 3]
        blank = "\n" if i % 11 == 0 else ""
        recs.append((name, seq, ">%s%s\n%s\n%s" % (name, comment, body, blank)))
    text = "".join(r[2] for r in recs)
    for tail in ("", "no-final-newline"):
        data = text if not tail else text[:-1] if text.endswith("\n") and not text.endswith("\n\n") else text
        path = tmp_path / ("d%s.fa" % tail)
        path.write_bytes(data.encode("latin1"))
        got = _fast_dump(path, threads)
        assert got[-1] == "END -1" and len(got) == len(recs) + 1
        ref = subprocess.run([DUMP, str(path)], stdout=subprocess.PIPE, check=True).stdout.decode("latin1").split("\n")[:-1]
        for g, r, (name, seq, _) in zip(got, ref, recs):
            n, gname, gseq, head, tail_b = g.split("\t")
            rl, rname, _, rseq, _ = r.split("\t")
            assert (int(n), gname, gseq) == (int(rl), rname, rseq) == (len(seq), name, seq)
            cut = min(7, len(seq) // 2)
            assert head == seq[:cut] and tail_b == seq[len(seq) - cut:]


@pytest.mark.parametrize("data", [b"@r1\nACGT\n+\nIIII\n", b">a\nAC\r\nGT\n", b">a\nACGT\n+\nIIII\n", b">a\nAC\n@b\nGT\n",
                                  b">a\nAC\x00GT\n", b"junk\n>a\nACGT\n", b">a\nAC\n>b x\r\nGT\n"])
def test_fast_fasta_declines_anything_but_strict_plain_fasta(data, tmp_path):
    _build()
    path = tmp_path / "d.fa"
    path.write_bytes(data)
    assert _fast_dump(path, 2) == ["FALLBACK"]
    import gzip as _gz
    with _gz.open(tmp_path / "d.fa.gz", "wb") as f:
        f.write(b">a\nACGT\n")
    assert _fast_dump(tmp_path / "d.fa.gz", 2) == ["FALLBACK"]


# ---- block-parallel ingest (arcs_b200/host/ingest.h) against the faithful sequential record loop ----
INGEST = os.path.join(ROOT, "arcs_b200", "bin", "ingest_dump")


def _build_ingest():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "arcs_b200", "host"), "../bin/ingest_dump"])


def _ingest(mode, path, *extra):
    p = subprocess.run([INGEST, mode, str(path)] + [str(x) for x in extra], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    fast = int(p.stderr.decode().split("FAST_BLOCKS")[1].split()[0])
    return p.stdout, fast


def _strict_fastq(rng, n_pairs, barcodes=7, tweak=None):
    recs = []
    for i in range(n_pairs):
        bc = "BX:Z:%s-1" % "".join("ACGT"[(int(rng.integers(0, barcodes)) >> (2 * q)) & 3] for q in range(6))
        for m in (1, 2):
            L = int(rng.integers(20, 160))
            seq = "".join("ACGTNacgt"[int(x)] for x in rng.integers(0, 9, L))
            recs.append(["r%d/%d" % (i, m), bc, seq, "I" * L])
    if tweak:
        tweak(recs)
    return "".join("@%s%s\n%s\n+\n%s\n" % (n, (" " + c) if c is not None else "", s, q) for n, c, s, q in recs).encode()


def _tweak_pairing(recs):
    recs[4][0] = "other/1"          # unpaired names (message + counter)
    recs[10][1] = None              # no comment
    recs[13][1] = "RG:Z:x"          # no BX on mate 2
    recs[16][1] = recs[16][1][:-1] + "2"  # barcodes differ
    recs[20][1] = "XY:Z:1 " + recs[20][1] + " ZZ:i:3"
    recs[21][1] = "XY:Z:1 " + recs[21][1] + " ZZ:i:3"
    recs[24][1] = "BX:Z:"           # empty barcode: counted under the empty name
    recs[25][1] = "BX:Z:"
    recs[30][0] = recs[30][0].replace("/1", "/1x")
    recs[32][1] = recs[32][1] + "\tQT:Z:x"  # a tab does not end the barcode


@pytest.mark.parametrize("gz", [False, True])
@pytest.mark.parametrize("workers,block", [(1, 2000), (3, 1000), (4, 1 << 16), (8, 1 << 20)])
def test_parallel_ingest_equals_sequential_on_strict_fastq(workers, block, gz, tmp_path):
    _build_ingest()
    rng = np.random.default_rng(workers * 1000 + block)
    data = _strict_fastq(rng, 300, tweak=_tweak_pairing)
    path = tmp_path / ("r.fq.gz" if gz else "r.fq")
    if gz:
        with gzip.open(path, "wb") as f:
            f.write(data)
    else:
        path.write_bytes(data)
    want, _ = _ingest("seq", path)
    got, fast = _ingest("par", path, workers, block)
    assert got == want
    assert b"File contains unpaired reads: other r2" in got
    if block < len(data):
        assert fast > 0  # the block path did run


@pytest.mark.parametrize("shards", [2, 3, 8])
def test_sharded_blocks_deal_every_barcode_to_one_gpu(shards, tmp_path):
    """--gpus N: the parser threads lay every block out by shard (barcode_shard of the barcode text); the pairs and
    their order inside a shard are those of the sequential loop, and a barcode never reaches two shards."""
    _build_ingest()
    rng = np.random.default_rng(shards)
    data = _strict_fastq(rng, 400, barcodes=40, tweak=_tweak_pairing)
    path = tmp_path / "r.fq"
    path.write_bytes(data)

    def run(mode, *extra):
        p = subprocess.run([INGEST, mode, str(path)] + [str(x) for x in extra], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           check=True, env=dict(os.environ, ARKS_SHARDS=str(shards)))
        return p.stdout.decode().split("\n")

    want = run("seq")
    got = run("par", 3, 3000)
    per_shard = lambda lines, g: [l for l in lines if l.startswith("P\t") and l.endswith("\t%d" % g)]  # noqa: E731
    seen = {}
    for g in range(shards):
        assert per_shard(got, g) == per_shard(want, g)
        for l in per_shard(got, g):
            assert seen.setdefault(l.split("\t")[1], g) == g
    assert len(set(seen.values())) > 1
    assert [l for l in got if not l.startswith("P\t")] == [l for l in want if not l.startswith("P\t")]
    # one shard: the output carries no shard column and equals the unsharded run
    plain = subprocess.run([INGEST, "par", str(path), "3", "3000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout.decode()
    assert sorted(l.rsplit("\t", 1)[0] for l in got if l.startswith("P\t")) == sorted(l for l in plain.split("\n") if l.startswith("P\t"))


# base = strict text, c = offset of a record-pair boundary in it
IRREGULAR = {
    "crlf_in_the_middle": lambda d, c: d[:c] + d[c:].replace(b"\n", b"\r\n", 8),
    "multiline_record": lambda d, c: d[:c] + b"@ml/1 BX:Z:AAAAAA-1\nACGT\nACGT\n+\nIIII\nIIII\n@ml/2 BX:Z:AAAAAA-1\nACGTACGT\n+\nIIIIIIII\n" + d[c:],
    "fasta_record_inside": lambda d, c: d[:c] + b">f1 BX:Z:AAAAAA-1\nACGTACGT\n>f2 BX:Z:AAAAAA-1\nACGTACGT\n" + d[c:],
    "nul_byte": lambda d, c: d[:c] + d[c:].replace(b"A", b"\0", 1),
    "empty_sequence": lambda d, c: d[:c] + b"@e/1 BX:Z:AAAAAA-1\n\n+\n\n@e/2 BX:Z:AAAAAA-1\n\n+\n\n" + d[c:],
    "odd_record_count": lambda d, c: d + b"@tail/1 BX:Z:AAAAAA-1\nACGT\n+\nIIII\n",
    "no_final_newline": lambda d, c: d[:-1],
    "truncated_quality": lambda d, c: d[:-30],
    "junk_between_records": lambda d, c: d[:c] + b"garbage line\n\n" + d[c:],
    "quality_longer_than_sequence": lambda d, c: d[:c] + b"@q/1 BX:Z:AAAAAA-1\nACGT\n+\nIIIIII\n" + d[c:],
    "one_extra_record": lambda d, c: d[:c] + b"@x/1 BX:Z:AAAAAA-1\nACGT\n+\nIIII\n" + d[c:],
    "not_fastq_at_all": lambda d, c: b">c1\nACGT\n>c2\nGGTA\n" * 50,
    "empty_file": lambda d, c: b"",
}


@pytest.mark.parametrize("name", sorted(IRREGULAR))
@pytest.mark.parametrize("workers,block", [(2, 1500), (4, 4096)])
def test_parallel_ingest_falls_back_to_the_faithful_reader(name, workers, block, tmp_path):
    """wherever the input stops being strict 4-line FASTQ the rest goes through the sequential reader: same
    pairs, counters, messages and multiplicities (incl. readBarcodes' stop at the first empty sequence)"""
    _build_ingest()
    rng = np.random.default_rng(5)
    base = _strict_fastq(rng, 60)
    lines = base.split(b"\n")
    n = base[:len(base) // 2].count(b"\n") // 8 * 8
    cut = len(b"\n".join(lines[:n])) + 1
    assert base[cut:cut + 1] == b"@" and base[cut - 1:cut] == b"\n"
    data = IRREGULAR[name](base, cut)
    path = tmp_path / "r.fq"
    path.write_bytes(data)
    want, _ = _ingest("seq", path)
    got, _ = _ingest("par", path, workers, block)
    assert got == want


def test_parallel_ingest_with_known_multiplicities(tmp_path):
    """-u: barcodes missing from the multiplicity file are rejected inside the block parser"""
    _build_ingest()
    rng = np.random.default_rng(9)
    data = _strict_fastq(rng, 400, barcodes=12)
    path = tmp_path / "r.fq"
    path.write_bytes(data)
    seen = sorted({ln.split(b"BX:Z:")[1] for ln in data.split(b"\n") if ln.startswith(b"@")})
    mult = tmp_path / "m.csv"
    mult.write_text("".join("%s,%d\n" % (b.decode(), 60 + i if i % 4 else 3) for i, b in enumerate(seen) if i % 3))
    want, _ = _ingest("seq", path, 0, 0, mult)
    got, fast = _ingest("par", path, 3, 2000, mult)
    assert got == want and fast > 0
    assert b"invalidbarcode=0" not in got.split(b"COUNTERS")[1].split(b"\n")[0]


def test_parse_bx_tag_known_answers():
    """the reference's own unit test of the SAM tag parser (Test/SAMTest.cpp:8-15), applied to the FASTQ-comment
    rule of ingest.h (first BX:Z: tag up to the next blank) -- on these vectors the two rules agree"""
    _build_ingest()
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        fq = os.path.join(tmp, "t.fq")
        with open(fq, "w") as f:
            f.write("@r1/1 QT:Z:AA<FFKKK BX:Z:CGTCAGGTCAGAGGTG-1 XT:i:0\nACGT\n+\nIIII\n@r1/2 QT:Z:AA<FFKKK BX:Z:CGTCAGGTCAGAGGTG-1 XT:i:0\nACGT\n+\nIIII\n"
                    "@r2/1 QT:Z:AA<FFKKK XT:i:0\nACGT\n+\nIIII\n@r2/2 QT:Z:AA<FFKKK XT:i:0\nACGT\n+\nIIII\n")
        out, _ = _ingest("seq", fq)
        assert out.split(b"\n")[0] == b"P\tCGTCAGGTCAGAGGTG-1\tACGT\tACGT"
        assert b"emptybarcode=1" in out
        assert _ingest("par", fq, 2, 64)[0] == out


# ---- fast_inflate.h (the gzip decoder of the read ingest) against zlib ------------------------------------
INFLATE = os.path.join(ROOT, "arcs_b200", "bin", "inflate_check")


def _build_inflate():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "arcs_b200", "host"), "../bin/inflate_check"])


def _fast_inflate(path, chunk=1 << 20):
    p = subprocess.run([INFLATE, "fast", str(path), str(chunk)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    err = [ln for ln in p.stderr.decode().splitlines() if ln.startswith("ERROR")]
    return p.stdout, (err[0] if err else None)


def _gz(data, level=6, strategy=0, mem_level=8):
    import zlib
    c = zlib.compressobj(level, zlib.DEFLATED, 31, mem_level, strategy)
    return c.compress(data) + c.flush()


def _payloads():
    rng = np.random.default_rng(3)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    fastq = b"".join(b"@r%d BX:Z:ACGTACGTAC-1\n%s\n+\n%s\n" % (i, acgt[rng.integers(0, 4, 150)].tobytes(),
                                                          bytes(rng.integers(35, 75, 150, dtype=np.uint8))) for i in range(4000))
    far = bytes(rng.integers(0, 256, 32768, dtype=np.uint8))
    return {
        "empty": b"",
        "one_byte": b"x",
        "fastq": fastq,
        "random_bytes": bytes(rng.integers(0, 256, 300_000, dtype=np.uint8)),
        "runs": b"A" * 100_000 + b"AB" * 50_000 + b"ABC" * 40_000 + b"ABCDEFG" * 20_000,
        "window_edge": (far + b"-" + far[:20000] + far[1:] + b"+" + far) * 3,  # matches at distance 32768 and 32767
        "text": b" ".join(b"w%d" % (i * 7919 % 1000) for i in range(200_000)),
    }


@pytest.mark.parametrize("name", sorted(_payloads()))
def test_fast_inflate_matches_zlib_on_every_block_type(name, tmp_path):
    import zlib
    _build_inflate()
    data = _payloads()[name]
    variants = [(lv, 0, 8) for lv in (0, 1, 4, 6, 9)] + [(6, zlib.Z_FIXED, 8), (6, zlib.Z_HUFFMAN_ONLY, 8), (6, zlib.Z_RLE, 8),
                                                           (6, zlib.Z_FILTERED, 8), (9, 0, 1), (1, 0, 9)]
    for i, (lv, strat, ml) in enumerate(variants):
        path = tmp_path / ("v%d.gz" % i)
        path.write_bytes(_gz(data, lv, strat, ml))
        out, err = _fast_inflate(path, chunk=(1 << 20) if i % 2 else 4099)
        assert err is None, (name, lv, strat, err)
        assert out == data, (name, lv, strat)


def test_fast_inflate_members_headers_and_chunk_sizes(tmp_path):
    import struct
    import zlib
    _build_inflate()
    p = _payloads()
    # several members, an empty one among them, then trailing zeros (ignored like zlib does)
    blob = _gz(p["fastq"][:50000]) + _gz(b"") + _gz(p["text"][:70001], 9) + _gz(p["runs"], 1)
    want = p["fastq"][:50000] + p["text"][:70001] + p["runs"]
    (tmp_path / "multi.gz").write_bytes(blob + b"\0" * 37)
    for chunk in (1, 7, 4096, 1 << 20):
        if chunk == 1 and len(want) > 300000:
            continue
        out, err = _fast_inflate(tmp_path / "multi.gz", chunk)
        assert err is None and out == want
    # header with FEXTRA, FNAME, FCOMMENT and FHCRC
    raw = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = raw.compress(p["text"]) + raw.flush()
    hdr = b"\x1f\x8b\x08" + bytes([4 | 8 | 16 | 2]) + b"\0\0\0\0\x00\x03" + struct.pack("<H", 5) + b"EXTRA" + b"name.fq\0" + b"a comment\0"
    hdr += struct.pack("<H", zlib.crc32(hdr) & 0xFFFF)
    (tmp_path / "flags.gz").write_bytes(hdr + body + struct.pack("<II", zlib.crc32(p["text"]), len(p["text"]) & 0xFFFFFFFF))
    out, err = _fast_inflate(tmp_path / "flags.gz")
    assert err is None and out == p["text"]
    # an empty file is an empty stream; a file that is not gzip is an error
    (tmp_path / "empty.gz").write_bytes(b"")
    assert _fast_inflate(tmp_path / "empty.gz") == (b"", None)
    (tmp_path / "plain.gz").write_bytes(b"@r1\nACGT\n+\nIIII\n" * 10)
    out, err = _fast_inflate(tmp_path / "plain.gz")
    assert out == b"" and err is not None


def test_fast_inflate_reports_truncation_and_corruption(tmp_path):
    _build_inflate()
    data = _payloads()["fastq"]
    blob = _gz(data)
    for cut in (len(blob) - 1, len(blob) - 8, len(blob) - 9, len(blob) // 2, 40, 12, 5):
        (tmp_path / "t.gz").write_bytes(blob[:cut])
        out, err = _fast_inflate(tmp_path / "t.gz")
        assert err is not None, cut
        assert data.startswith(out)  # everything handed out before the defect is right
    rng = np.random.default_rng(11)
    detected = 0
    for t in range(40):
        b = bytearray(blob)
        pos = int(rng.integers(20, len(b) - 8))
        b[pos] ^= 1 << int(rng.integers(0, 8))
        (tmp_path / "c.gz").write_bytes(bytes(b))
        out, err = _fast_inflate(tmp_path / "c.gz")
        detected += err is not None
        assert err is not None or out == data  # a flipped bit is either caught (structure / CRC-32 / length) or harmless
    assert detected >= 38
    # wrong CRC and wrong ISIZE in the trailer
    for off in (-8, -4):
        b = bytearray(blob)
        b[off] ^= 0x55
        (tmp_path / "c.gz").write_bytes(bytes(b))
        assert _fast_inflate(tmp_path / "c.gz")[1] is not None


# ---- par_inflate.h (multi-threaded decoding of one gzip member) against the data and against zlib ----------
def _par_inflate(path, threads=4, par_chunk=1 << 16, chunk=1 << 20):
    p = subprocess.run([INFLATE, "par%d" % threads, str(path), str(chunk)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True,
                       env=dict(os.environ, PAR_CHUNK=str(par_chunk)))
    lines = p.stderr.decode().splitlines()
    err = [ln for ln in lines if ln.startswith("ERROR")]
    nchunks = [int(ln.split()[1]) for ln in lines if ln.startswith("PARALLEL_CHUNKS")]
    return p.stdout, (err[0] if err else None), (nchunks[0] if nchunks else 0)


def _big_fastq(n_reads, seed):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGTN", dtype=np.uint8)
    recs = []
    for i in range(n_reads):
        L = int(rng.integers(100, 151))
        recs.append(b"@read%d/%d BX:Z:%s-1\n%s\n+\n%s\n" % (i // 2, i % 2 + 1, acgt[rng.integers(0, 4, 16)].tobytes(),
                                                           acgt[rng.integers(0, 5, L) % 4].tobytes(), bytes(rng.integers(35, 75, L, dtype=np.uint8))))
    return b"".join(recs)


@pytest.mark.parametrize("level", [1, 6, 9])
@pytest.mark.parametrize("threads", [1, 3, 8])
def test_par_inflate_fastq(level, threads, tmp_path):
    _build_inflate()
    data = _big_fastq(30000, level * 10 + threads)  # ~10 MB
    path = tmp_path / "r.fq.gz"
    path.write_bytes(_gz(data, level))
    out, err, nchunks = _par_inflate(path, threads, par_chunk=1 << 17)
    assert err is None and out == data
    assert nchunks >= 8  # the chunks really were decoded independently
    out, err, _ = _par_inflate(path, threads, par_chunk=1 << 16, chunk=4099)
    assert err is None and out == data


@pytest.mark.parametrize("name", sorted(_payloads()))
def test_par_inflate_any_payload(name, tmp_path):
    """whatever the data (binary data has no findable block starts, runs and far matches stress the window
    symbols), the result is the sequential one"""
    import zlib
    _build_inflate()
    data = _payloads()[name]
    for i, (lv, strat) in enumerate([(1, 0), (6, 0), (9, 0), (0, 0), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)]):
        path = tmp_path / ("v%d.gz" % i)
        path.write_bytes(_gz(data, lv, strat))
        out, err, _ = _par_inflate(path, 4, par_chunk=1 << 16)
        assert err is None, (name, lv, strat, err)
        assert out == data, (name, lv, strat)


def test_par_inflate_members_garbage_truncation_corruption(tmp_path):
    _build_inflate()
    a, b = _big_fastq(8000, 1), _big_fastq(3000, 2)
    # a second member and trailing zeros behind the first: the first goes parallel, the rest sequentially
    (tmp_path / "m.gz").write_bytes(_gz(a, 6) + _gz(b, 1) + b"\0" * 11)
    out, err, nchunks = _par_inflate(tmp_path / "m.gz", 4)
    assert err is None and out == a + b and nchunks > 4
    (tmp_path / "g.gz").write_bytes(_gz(a, 6) + b"\0" * 100)
    out, err, _ = _par_inflate(tmp_path / "g.gz", 4)
    assert err is None and out == a
    blob = _gz(a, 6)
    for cut in (len(blob) - 1, len(blob) - 9, len(blob) // 2, len(blob) // 3, 40):
        (tmp_path / "t.gz").write_bytes(blob[:cut])
        out, err, _ = _par_inflate(tmp_path / "t.gz", 4)
        assert err is not None, cut
        assert a.startswith(out), cut
    rng = np.random.default_rng(4)
    for t in range(30):
        c = bytearray(blob)
        c[int(rng.integers(20, len(c) - 8))] ^= 1 << int(rng.integers(0, 8))
        (tmp_path / "c.gz").write_bytes(bytes(c))
        out, err, _ = _par_inflate(tmp_path / "c.gz", 4)
        assert err is not None or out == a


# ---- bgzf_inflate.h (bgzip's many small members, dealt out to threads) ---------------------------------------
def _bgzf_member(chunk, level=6, extra_subfields=b""):
    import struct
    import zlib
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    d = c.compress(chunk) + c.flush()
    xlen = 6 + len(extra_subfields)
    bsize = 12 + xlen + len(d) + 8
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff" + struct.pack("<H", xlen) + extra_subfields + b"BC\x02\0" + struct.pack("<H", bsize - 1) + d +
            struct.pack("<II", zlib.crc32(chunk), len(chunk)))


def _bgzf(data, block=65280, level=6, eof=True, **kw):
    out = [_bgzf_member(data[i:i + block], level, **kw) for i in range(0, len(data), block)]
    if eof:
        out.append(_bgzf_member(b""))
    return b"".join(out)


def _bgzf_inflate(path, threads=4, wave=1 << 16, chunk=1 << 20):
    p = subprocess.run([INFLATE, "bgzf%d" % threads, str(path), str(chunk)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True,
                       env=dict(os.environ, PAR_CHUNK=str(wave)))
    lines = p.stderr.decode().splitlines()
    err = [ln for ln in lines if ln.startswith("ERROR")]
    members = [int(ln.split()[1]) for ln in lines if ln.startswith("PARALLEL_MEMBERS")]
    is_bgzf = [int(ln.split()[1]) for ln in lines if ln.startswith("IS_BGZF")]
    return p.stdout, (err[0] if err else None), (members[0] if members else 0), bool(is_bgzf and is_bgzf[0])


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_bgzf_inflate_fastq(threads, tmp_path):
    _build_inflate()
    data = _big_fastq(20000, 50 + threads)
    path = tmp_path / "r.fq.gz"
    blob = _bgzf(data, level=threads)
    assert gzip.decompress(blob) == data  # the writer above makes valid gzip
    path.write_bytes(blob)
    out, err, members, is_bgzf = _bgzf_inflate(path, threads, wave=1 << 18)
    assert is_bgzf and err is None and out == data
    assert members == (len(data) + 65279) // 65280 + 1
    out, err, _, _ = _bgzf_inflate(path, threads, wave=1 << 16, chunk=4099)
    assert err is None and out == data


@pytest.mark.parametrize("name", sorted(_payloads()))
def test_bgzf_inflate_any_payload(name, tmp_path):
    _build_inflate()
    data = _payloads()[name]
    for i, (block, level, eof) in enumerate([(65280, 6, True), (65000, 0, True), (1000, 1, False), (7, 9, True)]):
        if block < 100:
            data = data[:5000]
        path = tmp_path / ("v%d.gz" % i)
        path.write_bytes(_bgzf(data, block, level, eof))
        out, err, _, _ = _bgzf_inflate(path, 4)
        assert err is None, (name, block, level, err)
        assert out == data, (name, block, level)
    # other extra sub-fields in front of BC are skipped
    (tmp_path / "x.gz").write_bytes(_bgzf(data, 3000, extra_subfields=b"XY\x03\0abc"))
    out, err, members, is_bgzf = _bgzf_inflate(tmp_path / "x.gz", 4)
    assert is_bgzf and err is None and out == data and members > 0


def test_bgzf_inflate_mixed_members_garbage_truncation_corruption(tmp_path):
    _build_inflate()
    a, b, c = _big_fastq(6000, 11), _big_fastq(2000, 12), _big_fastq(1000, 13)
    # concatenated bgzf files (an empty end-of-file member in the middle), then a plain gzip member, then bgzf again
    # (sequential by then), then padding
    (tmp_path / "m.gz").write_bytes(_bgzf(a) + _bgzf(b, 20000) + _gz(c, 6) + _bgzf(a[:100000]) + b"\0" * 13)
    out, err, members, is_bgzf = _bgzf_inflate(tmp_path / "m.gz", 4)
    assert is_bgzf and err is None and out == a + b + c + a[:100000]
    assert members == (len(a) + 65279) // 65280 + 1 + (len(b) + 19999) // 20000 + 1
    # a plain gzip file is not bgzf; the decoder still reads it (all of it through the sequential decoder)
    (tmp_path / "p.gz").write_bytes(_gz(b, 6))
    out, err, members, is_bgzf = _bgzf_inflate(tmp_path / "p.gz", 4)
    assert not is_bgzf and err is None and out == b and members == 0
    blob = _bgzf(a)
    for cut in (len(blob) - 1, len(blob) - 29, len(blob) // 2, len(blob) // 3, 40, 17):
        (tmp_path / "t.gz").write_bytes(blob[:cut])
        out, err, _, _ = _bgzf_inflate(tmp_path / "t.gz", 4)
        assert a.startswith(out), cut
        assert err is not None, cut
    (tmp_path / "t.gz").write_bytes(blob[:-28])  # without the end-of-file member: still every byte
    out, err, _, _ = _bgzf_inflate(tmp_path / "t.gz", 4)
    assert err is None and out == a
    rng = np.random.default_rng(4)
    for t in range(40):
        d = bytearray(blob)
        d[int(rng.integers(0, len(d)))] ^= 1 << int(rng.integers(0, 8))
        (tmp_path / "c.gz").write_bytes(bytes(d))
        out, err, _, _ = _bgzf_inflate(tmp_path / "c.gz", 4)
        assert err is not None or out == a, t
        if err is not None:
            # what was delivered in front of the damage is right, except where the damage went to the sequential
            # decoder in the middle of a member (a broken header): then a prefix plus that decoder's bytes
            assert a.startswith(out[:(len(out) // 65280) * 65280]), t


def test_ingest_through_the_bgzf_decoder(tmp_path):
    _build_ingest()
    rng = np.random.default_rng(32)
    data = _strict_fastq(rng, 12000, barcodes=40, tweak=_tweak_pairing)
    path = tmp_path / "reads.fq.gz"
    path.write_bytes(_bgzf(data))
    ref = subprocess.run([INGEST, "seq", str(path)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True,
                         env=dict(os.environ, ARKS_ZLIB="1")).stdout
    for threads in ("1", "4"):
        got = subprocess.run([INGEST, "par", str(path), "4", str(1 << 18)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True,
                             env=dict(os.environ, ARKS_GZ_THREADS=threads)).stdout
        assert got == ref, threads


def test_par_inflate_several_large_members(tmp_path):
    """`cat a.fq.gz b.fq.gz c.fq.gz`: every member of some size is decoded by all threads, not only the first"""
    _build_inflate()
    parts = [_big_fastq(9000, 21), _big_fastq(7000, 22), _big_fastq(300, 23), _big_fastq(8000, 24)]
    blobs = [_gz(parts[0], 6), _gz(parts[1], 1), _gz(parts[2], 9), _gz(parts[3], 4)]
    (tmp_path / "m.gz").write_bytes(b"".join(blobs))
    only_first = None
    for threads in (2, 4):
        out, err, nchunks = _par_inflate(tmp_path / "m.gz", threads, par_chunk=1 << 16)
        assert err is None and out == b"".join(parts)
        (tmp_path / "first.gz").write_bytes(blobs[0])
        only_first = _par_inflate(tmp_path / "first.gz", threads, par_chunk=1 << 16)[2]
        assert nchunks > 2 * only_first  # the later members went through the parallel decoder too
    # damage in a later member: everything in front of it is still delivered, and the error is reported
    bad = bytearray(b"".join(blobs))
    bad[len(blobs[0]) + len(blobs[1]) // 2] ^= 0x10
    (tmp_path / "bad.gz").write_bytes(bytes(bad))
    out, err, _ = _par_inflate(tmp_path / "bad.gz", 4, par_chunk=1 << 16)
    assert err is not None and out.startswith(parts[0]) and (parts[0] + parts[1]).startswith(out[:len(parts[0]) + 1000])
    # a second member that is cut off
    (tmp_path / "cut.gz").write_bytes(blobs[0] + blobs[1][:len(blobs[1]) // 2])
    out, err, _ = _par_inflate(tmp_path / "cut.gz", 4, par_chunk=1 << 16)
    assert err is not None and out.startswith(parts[0]) and (parts[0] + parts[1]).startswith(out)


def test_ingest_through_the_parallel_gzip_decoder(tmp_path):
    """the whole read ingest (block-parallel parser) behind the multi-threaded gzip decoder gives the pairs,
    counters and multiplicities of the sequential reader behind zlib"""
    _build_ingest()
    rng = np.random.default_rng(31)
    data = _strict_fastq(rng, 12000, barcodes=40, tweak=_tweak_pairing)  # ~4 MB of FASTQ
    path = tmp_path / "reads.fq.gz"
    with gzip.open(path, "wb", compresslevel=4) as f:
        f.write(data)
    ref = subprocess.run([INGEST, "seq", str(path)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True,
                         env=dict(os.environ, ARKS_ZLIB="1")).stdout
    for threads in ("1", "3", "8"):
        got = subprocess.run([INGEST, "par", str(path), "4", str(1 << 18)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True,
                             env=dict(os.environ, ARKS_GZ_THREADS=threads)).stdout
        assert got == ref, threads
