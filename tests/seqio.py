"""Small test-side readers for WELL-FORMED FASTA / 4-line FASTQ (the demo fixtures),
plus the host-side ARKS pair filter restated in Python for small cases
(chromiumRead, Arcs.cpp:1185-1265).  Test infrastructure only."""
import gzip

import numpy as np


def _open(path):
    return gzip.open(path, "rb") if str(path).endswith(".gz") else open(path, "rb")


def read_fasta(path):
    """-> list of (name, seq bytes); name = text up to first whitespace"""
    out, name, chunks = [], None, []
    with _open(path) as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if name is not None:
                    out.append((name, b"".join(chunks)))
                name, chunks = line[1:].split()[0].decode() if line[1:].split() else "", []
            elif line:
                chunks.append(line)
    if name is not None:
        out.append((name, b"".join(chunks)))
    return out


def read_fastq(path):
    """-> list of (name, comment, seq) for strictly 4-line FASTQ"""
    out = []
    with _open(path) as f:
        while True:
            h = f.readline()
            if not h:
                break
            s = f.readline().rstrip(b"\r\n").split(b"\0")[0]  # the reference keeps C strings: cut at a NUL
            f.readline()
            f.readline()
            h = h.rstrip(b"\r\n")[1:]
            parts = h.split(None, 1)
            name = parts[0].decode() if parts else ""
            comment = parts[1].decode() if len(parts) > 1 else ""
            out.append((name, comment, s))
    return out


def strip_read_num(name):
    """stripReadNum (Arcs.cpp:243-254)"""
    pos = name.rfind("/")
    if pos == -1 or pos == 0 or pos == len(name) - 1:
        return name
    if not name[pos + 1].isdigit():
        return name
    return name[:pos]


def bx(comment):
    """barcode after the first 'BX:Z:' up to the next space (Arcs.cpp:1227-1251)"""
    i = comment.find("BX:Z:")
    if i < 0:
        return ""
    j = comment.find(" ", i)
    return comment[i + 5:j] if j >= 0 else comment[i + 5:]


def multiplicities(records):
    """readBarcodes (Arcs.cpp:481-547) for well-formed input"""
    mult = {}
    for name, comment, seq in records:
        if len(seq) <= 0:
            break
        if not comment:
            continue
        if "BX:Z:" in comment:
            b = bx(comment)
            mult[b] = mult.get(b, 0) + 1
    return mult


def candidate_pairs(records, mult):
    """pairs that reach checkReadSequence/bestContig -> (barcode list, bases uint8, off uint32[2n+1])"""
    barcodes, seqs = [], []
    for i in range(0, len(records) - 1, 2):
        n1, c1, s1 = records[i]
        n2, c2, s2 = records[i + 1]
        if strip_read_num(n1) != strip_read_num(n2):
            continue
        b1, b2 = bx(c1), bx(c2)
        if not b1 or not b2 or b1 != b2 or b1 not in mult:
            continue
        barcodes.append(b1)
        seqs.append(s1)
        seqs.append(s2)
    off = np.zeros(len(seqs) + 1, dtype=np.uint32)
    if seqs:
        off[1:] = np.cumsum([len(s) for s in seqs])
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8) if seqs else np.zeros(0, dtype=np.uint8)
    return barcodes, bases, off
