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
