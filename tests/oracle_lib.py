"""ctypes binding of oracle/libarks_oracle.so (the CPU restatement) and of
oracle/_ref/libref_prepseq.so (the reference's unmodified ReadsProcessor).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg; never by arcs_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
REF_BIN = os.path.join(REF_DIR, "arcs_ref")


class IndexStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("kmers_valid", "kmers_null", "recorded", "collisions", "removed")] + [
        ("unique", C.c_int64)
    ]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class MapStats(C.Structure):
    _fields_ = [
        (n, C.c_uint64)
        for n in (
            "kmers_valid",
            "kmers_invalid",
            "found",
            "recorded",
            "dups",
            "reads_pass",
            "reads_fail",
            "pairs_stored",
            "pairs_invalid",
            "pairs_nogood",
        )
    ]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build_oracle():
    """(Re)build libarks_oracle.so if missing or stale."""
    so = os.path.join(ORACLE_DIR, "libarks_oracle.so")
    src = os.path.join(ORACLE_DIR, "arks_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "libarks_oracle.so"], stderr=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        u8p, i32p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)
        L.arks_oracle_key.argtypes = [C.c_char_p, C.c_int, u8p]
        L.arks_oracle_key.restype = C.c_int
        L.arks_oracle_kmap_new.argtypes = [C.c_int, C.c_uint64]
        L.arks_oracle_kmap_new.restype = C.c_void_p
        L.arks_oracle_kmap_free.argtypes = [C.c_void_p]
        L.arks_oracle_kmap_size.argtypes = [C.c_void_p]
        L.arks_oracle_kmap_size.restype = C.c_uint64
        L.arks_oracle_kmap_dump.argtypes = [C.c_void_p, u8p, i32p]
        L.arks_oracle_map_kmers.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(IndexStats)]
        L.arks_oracle_map_kmers.restype = C.c_int
        L.arks_oracle_end_cutoff.argtypes = [C.c_int, C.c_int]
        L.arks_oracle_end_cutoff.restype = C.c_int
        L.arks_oracle_check_read.argtypes = [C.c_char_p, C.c_int]
        L.arks_oracle_check_read.restype = C.c_int
        L.arks_oracle_best_contig.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_double, C.POINTER(MapStats)]
        L.arks_oracle_best_contig.restype = C.c_int
        L.arks_oracle_map_pairs.argtypes = [
            C.c_void_p,
            C.c_char_p,
            u32p,
            C.c_uint64,
            C.c_double,
            i32p,
            C.POINTER(MapStats),
        ]
        L.arks_oracle_normal_estimation.argtypes = [C.c_int, C.c_float, C.c_int]
        L.arks_oracle_normal_estimation.restype = C.c_float
        L.arks_oracle_head_or_tail.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float]
        L.arks_oracle_head_or_tail.restype = C.c_int
        L.arks_oracle_pair_contigs.argtypes = [u32p, u32p, u32p, u32p, C.c_uint64, i32p, C.c_int, C.c_int, C.c_int,
                                               C.c_float, u32p, u32p, u32p, u32p, C.c_uint64]
        L.arks_oracle_pair_contigs.restype = C.c_uint64
        L.arks_oracle_edge.argtypes = [u32p, C.c_int, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.arks_oracle_edge.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def key(window: bytes, k: int):
    """canonical key bytes of one window, or None (prepSeq's NULL)"""
    out = np.zeros((k + 3) // 4, dtype=np.uint8)
    ok = lib().arks_oracle_key(window, k, _p(out, C.c_uint8))
    return bytes(out) if ok else None


class KMap:
    """exact kmer -> conreci map built with the reference's insert rule"""

    def __init__(self, k, expected=1024):
        self.k = k
        self.nb = (k + 3) // 4
        self.h = lib().arks_oracle_kmap_new(k, expected)
        self.stats = IndexStats()

    def __del__(self):
        if getattr(self, "h", None):
            lib().arks_oracle_kmap_free(self.h)
            self.h = None

    def __len__(self):
        return int(lib().arks_oracle_kmap_size(self.h))

    def map_kmers(self, seq: bytes, conreci: int):
        return lib().arks_oracle_map_kmers(self.h, seq, len(seq), conreci, C.byref(self.stats))

    def add_contig(self, seq: bytes, first_conreci: int, end_length: int):
        """getContigKmers body for one contig with len >= min_size (Arcs.cpp:1056-1094)"""
        cut = lib().arks_oracle_end_cutoff(len(seq), end_length)
        self.map_kmers(seq[:cut], first_conreci)
        self.map_kmers(seq[len(seq) - cut:], first_conreci + 1)

    def dump(self):
        n = len(self)
        keys = np.zeros((n, self.nb), dtype=np.uint8)
        vals = np.zeros(n, dtype=np.int32)
        if n:
            lib().arks_oracle_kmap_dump(self.h, _p(keys, C.c_uint8), _p(vals, C.c_int32))
        return keys, vals

    def best_contig(self, read: bytes, j: float, stats=None):
        st = stats if stats is not None else MapStats()
        return lib().arks_oracle_best_contig(self.h, read, len(read), j, C.byref(st))

    def map_pairs(self, bases: np.ndarray, off: np.ndarray, j: float):
        """bases: uint8 array; off: uint32[2n+1] -> (conreci int32[n], MapStats)"""
        n = (len(off) - 1) // 2
        out = np.zeros(n, dtype=np.int32)
        st = MapStats()
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint32)
        lib().arks_oracle_map_pairs(self.h, bases.ctypes.data_as(C.c_char_p), _p(off, C.c_uint32), n, j,
                                    _p(out, C.c_int32), C.byref(st))
        return out, st


def head_or_tail(head, tail, min_reads, error_percent):
    r = lib().arks_oracle_head_or_tail(head, tail, min_reads, error_percent)
    return bool(r & 1), bool(r & 2)


def pair_contigs(barcode, contig, head, tail, mult, min_mult, max_mult, min_reads, error_percent, rank):
    barcode = np.ascontiguousarray(barcode, dtype=np.uint32)
    contig = np.ascontiguousarray(contig, dtype=np.uint32)
    head = np.ascontiguousarray(head, dtype=np.uint32)
    tail = np.ascontiguousarray(tail, dtype=np.uint32)
    mult = np.ascontiguousarray(mult, dtype=np.int32)
    rank = np.ascontiguousarray(rank, dtype=np.uint32)
    cap = 1 << 16
    while True:
        a = np.zeros(cap, dtype=np.uint32)
        b = np.zeros(cap, dtype=np.uint32)
        c = np.zeros((cap, 4), dtype=np.uint32)
        n = lib().arks_oracle_pair_contigs(_p(barcode, C.c_uint32), _p(contig, C.c_uint32), _p(head, C.c_uint32),
                                           _p(tail, C.c_uint32), len(barcode), _p(mult, C.c_int32), min_mult, max_mult,
                                           min_reads, error_percent, _p(rank, C.c_uint32), _p(a, C.c_uint32),
                                           _p(b, C.c_uint32), _p(c, C.c_uint32), cap)
        if n <= cap:
            return a[:n], b[:n], c[:n]
        cap = int(n)


def edge(counts, min_links, error_percent):
    c = np.ascontiguousarray(counts, dtype=np.uint32)
    o, w = C.c_int(), C.c_int()
    ok = lib().arks_oracle_edge(_p(c, C.c_uint32), min_links, error_percent, C.byref(o), C.byref(w))
    return bool(ok), o.value, w.value


# ---- the reference's own prepSeq (only where oracle/_ref was built) -------------------------

_ref = None


def ref_prepseq_lib():
    global _ref
    if _ref is None:
        path = os.path.join(REF_DIR, "libref_prepseq.so")
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        L.ref_prepseq_all.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8)]
        L.ref_prepseq_all.restype = C.c_int
        _ref = L
    return _ref


def ref_keys_all(seq: bytes, k: int):
    """(keys uint8[nwin, nb], valid uint8[nwin]) from the reference's ReadsProcessor"""
    L = ref_prepseq_lib()
    nwin = max(0, len(seq) - k + 1)
    nb = (k + 3) // 4
    keys = np.zeros((nwin, nb), dtype=np.uint8)
    valid = np.zeros(nwin, dtype=np.uint8)
    if nwin:
        L.ref_prepseq_all(seq, len(seq), k, _p(keys, C.c_uint8), _p(valid, C.c_uint8))
    return keys, valid
