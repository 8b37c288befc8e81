"""Size-independent properties of the CUDA hot path at bench scale (millions of read pairs, tens of millions of
contig k-mers -- sizes the CPU oracle cannot finish in seconds): counter identities that hold for any input,
invariance under the batch split, additivity when the same reads are mapped twice, and agreement of the
lane-per-read kernel with the general warp-per-pair kernel (two independent implementations of bestContig on
the device).  Workload = bench.py's generators (BASELINE.json configs[1] shape), on the device."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

K, J, READ_LEN = 60, 0.55, 150
GENOME, CONTIGS, PAIRS, PPB = 10_000_000, 1000, 1_000_000, 250


def _workload():
    import torch

    import bench
    dev = torch.device("cuda", 0)
    genome, starts, ends = bench.make_draft(torch, dev, GENOME, CONTIGS, seed=1)
    iv = bench.contig_ends(starts, ends)
    end_bases = torch.cat([genome[s:e] for s, e, _ in iv])
    h_end_off = np.zeros(len(iv) + 1, dtype=np.uint64)
    h_end_off[1:] = np.cumsum([e - s for s, e, _ in iv])
    d_end_off = torch.from_numpy(h_end_off.astype(np.int64)).to(dev)
    d_conreci = torch.tensor([c for _, _, c in iv], dtype=torch.int32, device=dev)
    cfg = dict(bench.CONFIGS["c2"], ppb=PPB)
    bases, barcode, _ = bench.make_reads(torch, dev, genome, cfg, PAIRS, seed=2)
    return torch, dev, end_bases, h_end_off, d_end_off, d_conreci, bases, barcode


def _index(w, env=None):
    import arcs_b200
    torch, dev, end_bases, h_end_off, d_end_off, d_conreci, bases, barcode = w
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        idx = arcs_b200.ArksIndex(K, int(h_end_off[-1]), device=0)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v
    idx.set_stream(torch.cuda.current_stream().cuda_stream)
    idx.add_ends_device(end_bases.data_ptr(), d_end_off.data_ptr(), d_conreci.data_ptr(), h_end_off)
    idx.finalize()
    return idx


def _map(w, idx, splits):
    torch, dev, _, _, _, _, bases, barcode = w
    a = 0
    for n in splits:
        off = (torch.arange(2 * n + 1, device=dev, dtype=torch.int64) * READ_LEN).to(torch.int32)
        idx.map_pairs_device(bases.data_ptr() + a * 2 * READ_LEN, off.data_ptr(), barcode.data_ptr() + 4 * a, n, n * 2 * READ_LEN, J, None)
        torch.cuda.synchronize()
        a += n
    assert a == PAIRS


def _imap_sorted(idx):
    b, c, h, t = idx.imap()
    o = np.lexsort((c, b))
    return np.stack([b[o], c[o], h[o], t[o]], axis=1)


def test_full_size_properties():
    w = _workload()
    # one batch, lane-per-read kernel
    i1 = _index(w)
    _map(w, i1, [PAIRS])
    s1 = i1.map_stats().as_dict()
    m1 = _imap_sorted(i1)
    # counter identities (Arcs.cpp:957-1012,1266-1292): every window of a valid pair is valid or invalid; found = recorded + dups;
    # every read of a valid pair passes or fails the Jaccard gate; every pair is stored or not
    assert s1["kmers_valid"] + s1["kmers_invalid"] == 2 * (PAIRS - s1["pairs_invalid"]) * (READ_LEN - K + 1)
    assert 0 < s1["pairs_invalid"] < PAIRS // 100  # pairs with more than 2 % N never reach bestContig
    assert s1["found"] == s1["recorded"] + s1["dups"] and s1["found"] <= s1["kmers_valid"]
    assert s1["reads_pass"] + s1["reads_fail"] == 2 * (PAIRS - s1["pairs_invalid"])
    assert s1["pairs_stored"] + s1["pairs_nogood"] == PAIRS
    assert int(m1[:, 2].sum() + m1[:, 3].sum()) == s1["pairs_stored"]
    assert s1["pairs_stored"] > PAIRS // 2  # the workload maps
    # invariance under the batch split (ragged, incl. a batch of one pair and groups that straddle batches)
    i2 = _index(w)
    _map(w, i2, [1, 15, 333_333, 17, 500_000, PAIRS - 1 - 15 - 333_333 - 17 - 500_000])
    assert i2.map_stats().as_dict() == s1
    assert np.array_equal(_imap_sorted(i2), m1)
    # additivity: the same reads again double every tally and every counter
    _map(w, i2, [PAIRS])
    s2 = i2.map_stats().as_dict()
    assert all(s2[k] == 2 * s1[k] for k in s1)
    m2 = _imap_sorted(i2)
    assert np.array_equal(m2[:, :2], m1[:, :2]) and np.array_equal(m2[:, 2:], 2 * m1[:, 2:])
    del i2
    # the general warp-per-pair kernel (no lane path, no membership filter) gives the same answer
    i3 = _index(w, {"ARKS_MAP_MODE": "pair", "ARKS_BLOOM_BITS": "0"})
    _map(w, i3, [PAIRS])
    assert i3.map_stats().as_dict() == s1
    assert np.array_equal(_imap_sorted(i3), m1)
    del i3
    # so does the path without seed-and-extend at all (one table probe per window)
    i4 = _index(w, {"ARKS_NO_EXTEND": "1"})
    _map(w, i4, [PAIRS // 4, PAIRS - PAIRS // 4])
    assert i4.map_stats().as_dict() == s1
    assert np.array_equal(_imap_sorted(i4), m1)
