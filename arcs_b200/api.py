"""ctypes binding of libarks_b200.so (include/arks_b200.h), 1:1 with the C ABI.

`ArksIndex` mirrors the three seams of runArcs that the library replaces
(Arcs/Arcs.cpp:1871-1909): getContigKmers -> add_ends()/finalize(), readChroms ->
map_pairs(), pairContigs -> pair_links().  Inputs are numpy arrays in host memory, or
raw device pointers for the *_device variants (e.g. torch tensors' data_ptr()).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class ArksError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("arks error %d: %s" % (code, msg))
        self.code = code


class IndexStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("kmers_valid", "kmers_null", "recorded", "collisions", "removed", "unique")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class MapStats(C.Structure):
    _fields_ = [
        (n, C.c_uint64)
        for n in ("kmers_valid", "kmers_invalid", "found", "recorded", "dups", "reads_pass", "reads_fail",
                  "pairs_stored", "pairs_invalid", "pairs_nogood")
    ]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def lib_path():
    return os.path.join(_HERE, "lib", "libarks_b200.so")


_lib = None

# every symbol include/arks_b200.h declares
SYMBOLS = [
    "arks_create", "arks_destroy", "arks_last_error", "arks_set_stream", "arks_sync", "arks_host_alloc",
    "arks_host_free", "arks_index_add", "arks_index_add_device", "arks_index_finalize", "arks_index_size",
    "arks_index_dump", "arks_set_conreci_remap", "arks_map_pairs", "arks_map_pairs_device", "arks_map_get_stats",
    "arks_map_stats_reset", "arks_imap_size", "arks_imap_export", "arks_imap_add", "arks_pair_links",
    "arks_pmap_size", "arks_pmap_export", "arks_head_tail_table", "arks_launch_count", "arks_device_init",
    "arks_imap_clear", "arks_bind_thread", "arks_device_numa_node", "arks_map_pairs_begin", "arks_map_pairs_end", "arks_pmap_digest", "arks_comm_unique_id", "arks_comm_init_rank", "arks_comm_init_local", "arks_merge_pmap",
]


def load_library():
    """Loads libarks_b200.so; raises (no fallback) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ArksError(-1, "%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "or `make -C arcs_b200/csrc`" % path)
    L = C.CDLL(path)
    vp, u8p, i32p, u32p, u64p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    L.arks_create.argtypes = [C.c_int, C.c_int, C.c_uint64, C.POINTER(vp)]
    L.arks_destroy.argtypes = [vp]
    L.arks_destroy.restype = None
    L.arks_last_error.argtypes = [vp]
    L.arks_last_error.restype = C.c_char_p
    L.arks_set_stream.argtypes = [vp, vp]
    L.arks_sync.argtypes = [vp]
    L.arks_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.arks_host_free.argtypes = [vp]
    L.arks_device_init.argtypes = [C.c_int]
    L.arks_bind_thread.argtypes = [C.c_int]
    L.arks_device_numa_node.argtypes = [C.c_int, C.POINTER(C.c_int)]
    L.arks_index_add.argtypes = [vp, vp, u64p, u32p, C.c_uint32]
    L.arks_index_add_device.argtypes = [vp, vp, vp, vp, u64p, C.c_uint32]
    L.arks_index_finalize.argtypes = [vp, C.POINTER(IndexStats)]
    L.arks_index_size.argtypes = [vp, u64p]
    L.arks_index_dump.argtypes = [vp, u8p, i32p, C.c_uint64, u64p]
    L.arks_set_conreci_remap.argtypes = [vp, u32p, C.c_uint32]
    L.arks_map_pairs.argtypes = [vp, vp, u32p, u32p, C.c_uint32, C.c_double, i32p]
    L.arks_map_pairs_begin.argtypes = [vp, vp, u32p, u32p, C.c_uint32, C.c_double, i32p]
    L.arks_map_pairs_end.argtypes = [vp]
    L.arks_map_pairs_device.argtypes = [vp, vp, vp, vp, C.c_uint32, C.c_uint64, C.c_double, vp]
    L.arks_map_get_stats.argtypes = [vp, C.POINTER(MapStats)]
    L.arks_map_stats_reset.argtypes = [vp]
    L.arks_imap_size.argtypes = [vp, u64p]
    L.arks_imap_export.argtypes = [vp, u32p, u32p, u32p, u32p, C.c_uint64, u64p]
    L.arks_imap_clear.argtypes = [vp]
    L.arks_imap_add.argtypes = [vp, u32p, u32p, u32p, u32p, C.c_uint64]
    L.arks_pair_links.argtypes = [vp, i32p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_float, u32p, C.c_uint32]
    L.arks_pmap_size.argtypes = [vp, u64p]
    L.arks_pmap_export.argtypes = [vp, u32p, u32p, u32p, C.c_uint64, u64p]
    L.arks_head_tail_table.argtypes = [C.c_int, C.c_float, C.c_uint32, u32p]
    L.arks_pmap_digest.argtypes = [vp, u64p]
    L.arks_comm_unique_id.argtypes = [u8p]
    L.arks_comm_init_rank.argtypes = [vp, u8p, C.c_int, C.c_int]
    L.arks_comm_init_local.argtypes = [C.POINTER(vp), C.c_int]
    L.arks_merge_pmap.argtypes = [C.POINTER(vp), C.c_int]
    L.arks_launch_count.argtypes = [vp]
    L.arks_launch_count.restype = C.c_uint64
    for name in SYMBOLS:
        f = getattr(L, name)
        if name not in ("arks_destroy", "arks_last_error", "arks_launch_count"):
            f.restype = C.c_int
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def head_tail_table(min_reads, error_percent, n):
    """min_max[sum] decision table of headOrTail (Arcs.cpp:846-861); host-only, no GPU needed"""
    L = load_library()
    out = np.zeros(n, dtype=np.uint32)
    rc = L.arks_head_tail_table(int(min_reads), float(error_percent), n, _p(out, C.c_uint32))
    if rc:
        raise ArksError(rc, "head/tail predicate not monotone")
    return out


class ArksIndex:
    """One GPU's ARKS state: k-mer table + barcode tallies + pair links."""

    def __init__(self, k, max_kmers, device=0):
        self.L = load_library()
        self.k = k
        self.nb = (k + 3) // 4
        h = C.c_void_p()
        rc = self.L.arks_create(device, k, int(max_kmers), C.byref(h))
        if rc:
            raise ArksError(rc, self.L.arks_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.arks_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _ck(self, rc):
        if rc:
            raise ArksError(rc, self.L.arks_last_error(self.h).decode())

    # ---- plumbing
    def set_stream(self, cuda_stream):
        self._ck(self.L.arks_set_stream(self.h, C.c_void_p(cuda_stream)))

    def sync(self):
        self._ck(self.L.arks_sync(self.h))

    @property
    def launches(self):
        return int(self.L.arks_launch_count(self.h))

    # ---- kernel 1
    def add_ends(self, bases, end_off, conreci):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        end_off = np.ascontiguousarray(end_off, dtype=np.uint64)
        conreci = np.ascontiguousarray(conreci, dtype=np.uint32)
        assert len(end_off) == len(conreci) + 1
        self._ck(self.L.arks_index_add(self.h, bases.ctypes.data, _p(end_off, C.c_uint64), _p(conreci, C.c_uint32),
                                       len(conreci)))

    def add_ends_device(self, d_bases, d_end_off, d_conreci, h_end_off):
        h_end_off = np.ascontiguousarray(h_end_off, dtype=np.uint64)
        self._ck(self.L.arks_index_add_device(self.h, d_bases, d_end_off, d_conreci, _p(h_end_off, C.c_uint64),
                                              len(h_end_off) - 1))

    def finalize(self):
        st = IndexStats()
        self._ck(self.L.arks_index_finalize(self.h, C.byref(st)))
        return st

    def dump(self):
        n = C.c_uint64()
        self._ck(self.L.arks_index_size(self.h, C.byref(n)))
        keys = np.zeros((n.value, self.nb), dtype=np.uint8)
        vals = np.zeros(n.value, dtype=np.int32)
        if n.value:
            self._ck(self.L.arks_index_dump(self.h, _p(keys, C.c_uint8), _p(vals, C.c_int32), n.value, C.byref(n)))
        return keys, vals

    # ---- kernel 2
    def set_conreci_remap(self, remap):
        remap = np.ascontiguousarray(remap, dtype=np.uint32)
        self._ck(self.L.arks_set_conreci_remap(self.h, _p(remap, C.c_uint32), len(remap)))

    def map_pairs(self, bases, read_off, barcode_id, j_index, want_conreci=True):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        read_off = np.ascontiguousarray(read_off, dtype=np.uint32)
        barcode_id = np.ascontiguousarray(barcode_id, dtype=np.uint32)
        n = len(barcode_id)
        assert len(read_off) == 2 * n + 1
        out = np.zeros(n, dtype=np.int32) if want_conreci else None
        self._ck(self.L.arks_map_pairs(self.h, bases.ctypes.data, _p(read_off, C.c_uint32), _p(barcode_id, C.c_uint32),
                                       n, float(j_index), _p(out, C.c_int32) if want_conreci else None))
        return out

    def map_pairs_raw(self, bases_ptr, read_off_ptr, barcode_ptr, n_pairs, j_index):
        """host pointers (e.g. pinned buffers), no result copy: the end-to-end call"""
        self._ck(self.L.arks_map_pairs(self.h, bases_ptr, C.cast(read_off_ptr, C.POINTER(C.c_uint32)),
                                       C.cast(barcode_ptr, C.POINTER(C.c_uint32)), n_pairs, float(j_index), None))

    def map_pairs_device(self, d_bases, d_read_off, d_barcode_id, n_pairs, n_bases, j_index, d_conreci_out=None):
        self._ck(self.L.arks_map_pairs_device(self.h, d_bases, d_read_off, d_barcode_id, n_pairs, n_bases,
                                              float(j_index), d_conreci_out))

    def map_stats(self):
        st = MapStats()
        self._ck(self.L.arks_map_get_stats(self.h, C.byref(st)))
        return st

    def map_stats_reset(self):
        self._ck(self.L.arks_map_stats_reset(self.h))

    # ---- imap / pmap
    def imap(self):
        n = C.c_uint64()
        self._ck(self.L.arks_imap_size(self.h, C.byref(n)))
        a = [np.zeros(n.value, dtype=np.uint32) for _ in range(4)]
        if n.value:
            self._ck(self.L.arks_imap_export(self.h, *[_p(x, C.c_uint32) for x in a], n.value, C.byref(n)))
        return a  # barcode, contig, head, tail

    def imap_clear(self):
        self._ck(self.L.arks_imap_clear(self.h))

    def imap_add(self, barcode, contig, head, tail):
        a = [np.ascontiguousarray(x, dtype=np.uint32) for x in (barcode, contig, head, tail)]
        self._ck(self.L.arks_imap_add(self.h, *[_p(x, C.c_uint32) for x in a], len(a[0])))

    def pair_links_run(self, mult, min_mult, max_mult, min_reads, error_percent, lexrank):
        """pairContigs on the device; the ordered rows stay there (pmap_rows / pmap_digest / merge_pmap)"""
        mult = np.ascontiguousarray(mult, dtype=np.int32)
        lexrank = np.ascontiguousarray(lexrank, dtype=np.uint32)
        self._ck(self.L.arks_pair_links(self.h, _p(mult, C.c_int32), len(mult), min_mult, max_mult, min_reads,
                                        float(error_percent), _p(lexrank, C.c_uint32), len(lexrank)))

    def pmap_size(self):
        n = C.c_uint64()
        self._ck(self.L.arks_pmap_size(self.h, C.byref(n)))
        return int(n.value)

    def pmap_rows(self):
        n = self.pmap_size()
        a = np.zeros(n, dtype=np.uint32)
        b = np.zeros(n, dtype=np.uint32)
        c = np.zeros((n, 4), dtype=np.uint32)
        if n:
            nn = C.c_uint64()
            self._ck(self.L.arks_pmap_export(self.h, _p(a, C.c_uint32), _p(b, C.c_uint32), _p(c, C.c_uint32), n, C.byref(nn)))
        return a, b, c

    def pmap_export_raw(self, a_ptr, b_ptr, c_ptr, cap):
        """export into caller-owned host memory (pinned: arks_host_alloc) -> number of rows"""
        nn = C.c_uint64()
        u32p = C.POINTER(C.c_uint32)
        self._ck(self.L.arks_pmap_export(self.h, C.cast(a_ptr, u32p), C.cast(b_ptr, u32p), C.cast(c_ptr, u32p), cap, C.byref(nn)))
        return int(nn.value)

    def pmap_digest(self):
        d = (C.c_uint64 * 2)()
        self._ck(self.L.arks_pmap_digest(self.h, d))
        return int(d[0]), int(d[1])

    def pair_links(self, mult, min_mult, max_mult, min_reads, error_percent, lexrank):
        self.pair_links_run(mult, min_mult, max_mult, min_reads, error_percent, lexrank)
        return self.pmap_rows()

    # ---- multi-GPU (one process per GPU): the single NCCL exchange of the pair-link map
    def comm_init_rank(self, comm_id, rank, n_ranks):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(comm_id))
        self._ck(self.L.arks_comm_init_rank(self.h, buf, rank, n_ranks))

    def merge_pmap(self):
        hs = (C.c_void_p * 1)(self.h)
        self._ck(self.L.arks_merge_pmap(hs, 1))


def comm_unique_id():
    """ncclUniqueId (128 bytes) made by this process; hand it to every rank's comm_init_rank"""
    L = load_library()
    buf = (C.c_uint8 * 128)()
    rc = L.arks_comm_unique_id(buf)
    if rc:
        raise ArksError(rc, L.arks_last_error(None).decode())
    return bytes(buf)


def comm_init_local(indexes):
    """one process, several handles (one per GPU, or several shards on one GPU in tests)"""
    L = load_library()
    hs = (C.c_void_p * len(indexes))(*[i.h for i in indexes])
    rc = L.arks_comm_init_local(hs, len(indexes))
    if rc:
        raise ArksError(rc, L.arks_last_error(indexes[0].h).decode())


def merge_pmap_local(indexes):
    L = load_library()
    hs = (C.c_void_p * len(indexes))(*[i.h for i in indexes])
    rc = L.arks_merge_pmap(hs, len(indexes))
    if rc:
        raise ArksError(rc, L.arks_last_error(indexes[0].h).decode())
