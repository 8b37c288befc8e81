"""Multi-GPU path on the device: read pairs sharded by barcode over several handles, every shard's pair links
ordered on the device, one exchange (arks_merge_pmap: all-gather of keys + all-reduce of counters) -- the merged
map must equal the oracle's pairContigs on the whole input and the one-handle result, whatever the number of
shards.  Shards on ONE device exchange by device-to-device copies (runs on any box); N processes on N devices
exchange over NCCL (skipped when the box has fewer GPUs)."""
import os
import socket

import numpy as np
import pytest

import glue
import oracle_lib as O
from tools import synth

pytestmark = pytest.mark.gpu

K, J = 32, 0.4
LINK_ARGS = (40, 10000, 2, 0.05)  # min_mult, max_mult, min_reads, error_percent


def _workload():
    rng = np.random.default_rng(77)
    genome, contigs = synth.make_draft(rng, 400000, 5000, K)
    bases, end_off, conreci, names = synth.contig_end_arrays(genome, contigs, K, end_length=2000)
    rb, roff, bc = synth.make_reads(rng, genome, n_barcodes=240, pairs_per_barcode=120, mol_len=30000, mols_per_barcode=2)
    return bases, end_off, conreci, names, rb, roff, bc


def _index(A, bases, end_off, conreci, device=0):
    idx = A.ArksIndex(K, int(end_off[-1]) + 16, device=device)
    idx.add_ends(bases, end_off, conreci)
    idx.finalize()
    return idx


def _map_shard(idx, rb, roff, bc, keep):
    """maps the pairs whose barcode is in this shard (keep: bool per pair)"""
    sel = np.nonzero(keep)[0]
    if len(sel) == 0:
        return
    parts, offs, pos = [], [0], 0
    for i in sel:
        for r in (2 * i, 2 * i + 1):
            parts.append(rb[roff[r]:roff[r + 1]])
            pos += int(roff[r + 1] - roff[r])
            offs.append(pos)
    idx.map_pairs(np.concatenate(parts), np.array(offs, dtype=np.uint32), bc[sel], J, want_conreci=False)


def _oracle_links(bases, end_off, conreci, names, rb, roff, bc):
    km = O.KMap(K, int(end_off[-1]))
    for e in range(len(conreci)):
        km.map_kmers(bases[int(end_off[e]):int(end_off[e + 1])].tobytes(), int(conreci[e]))
    cr, _ = km.map_pairs(rb, roff, J)
    imap, uniq = glue.imap_rows([str(b) for b in bc], cr, names)
    rows = np.array(sorted((int(b), c, ht[0], ht[1]) for b, d in imap.items() for c, ht in d.items()), dtype=np.uint32)
    rank = glue.lex_rank(uniq)
    nb = int(bc.max()) + 1
    mult = np.bincount(bc, minlength=nb).astype(np.int32) * 2
    want = O.pair_contigs(rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3], mult, LINK_ARGS[0], LINK_ARGS[1], LINK_ARGS[2],
                          np.float32(LINK_ARGS[3]), rank)
    return want, mult, rank


@pytest.mark.parametrize("shards", [1, 2, 3, 5])
def test_shards_on_one_device_merge_to_the_oracle_map(shards, monkeypatch):
    import arcs_b200 as A
    bases, end_off, conreci, names, rb, roff, bc = _workload()
    (oa, ob, oc), mult, rank = _oracle_links(bases, end_off, conreci, names, rb, roff, bc)
    assert len(oa) > 100
    if shards == 3:
        monkeypatch.setenv("ARKS_PMAP_INITIAL_SLOTS", "64")  # the pmap hash has to grow (and the pass repeat) several times
    idxs = [_index(A, bases, end_off, conreci) for _ in range(shards)]
    for g, idx in enumerate(idxs):
        _map_shard(idx, rb, roff, bc, bc % shards == g)
        idx.pair_links_run(mult, LINK_ARGS[0], LINK_ARGS[1], LINK_ARGS[2], LINK_ARGS[3], rank)
    if shards > 1:
        sizes = [i.pmap_size() for i in idxs]
        assert max(sizes) < len(oa)  # no shard holds the whole map before the exchange
    A.comm_init_local(idxs)
    A.merge_pmap_local(idxs)
    digests = {i.pmap_digest() for i in idxs}
    assert len(digests) == 1
    for idx in idxs:
        a, b, c = idx.pmap_rows()
        assert np.array_equal(a, oa) and np.array_equal(b, ob) and np.array_equal(c, oc)
    # rows are in (rank a, rank b) order = std::map<pair<string,string>> iteration order
    keys = (rank[oa].astype(np.int64) << 32) | rank[ob]
    assert np.all(np.diff(keys) > 0)
    for idx in idxs:
        idx.close()


def _nccl_rank(rank_id, world, port, out):
    import torch
    import torch.distributed as dist

    import arcs_b200 as A
    from arcs_b200 import merge
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank_id)
    dev = torch.device("cuda", rank_id)
    dist.init_process_group("nccl", rank=rank_id, world_size=world, device_id=dev)
    bases, end_off, conreci, names, rb, roff, bc = _workload()
    _, uniq = glue.name_ids(names)
    rank = glue.lex_rank(uniq)
    mult = np.bincount(bc, minlength=int(bc.max()) + 1).astype(np.int32) * 2
    idx = _index(A, bases, end_off, conreci, device=rank_id)
    _map_shard(idx, rb, roff, bc, merge.shard_of_barcode(bc, world) == rank_id)
    idx.pair_links_run(mult, LINK_ARGS[0], LINK_ARGS[1], LINK_ARGS[2], LINK_ARGS[3], rank)
    merge.init_comm(idx, dev)
    merge.merge_pmap(idx)
    a, b, c = idx.pmap_rows()
    d = idx.pmap_digest()
    np.savez(out % rank_id, a=a, b=b, c=c, d=np.array(d, dtype=np.uint64))
    idx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_merge_over_real_devices(world, tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "rank%d.npz")
    mp.spawn(_nccl_rank, args=(world, port, out), nprocs=world, join=True)
    bases, end_off, conreci, names, rb, roff, bc = _workload()
    (oa, ob, oc), _, _ = _oracle_links(bases, end_off, conreci, names, rb, roff, bc)
    digests = set()
    for r in range(world):
        got = np.load(out % r)
        assert np.array_equal(got["a"], oa) and np.array_equal(got["b"], ob) and np.array_equal(got["c"], oc)
        digests.add(tuple(got["d"].tolist()))
    assert len(digests) == 1


def test_pair_links_with_millions_of_rows(monkeypatch):
    """the pair-link stage at scale: 2 000 barcodes x 110 contigs each -> 12 M pair events, ~4 M distinct pairs; the
    device's hash (grown twice from a deliberately small start), multi-tile radix sort and row gather against a numpy
    restatement of pairContigs' counting (Arcs.cpp:1384-1432) -- rows, order and all four orientation counters"""
    import arcs_b200 as A
    rng = np.random.default_rng(5)
    n_ct, n_bc, per = 3000, 2000, 110
    idx = A.ArksIndex(32, 1024)
    seq = synth.ACGT[rng.integers(0, 4, 600)]
    idx.add_ends(seq, np.array([0, 300, 600], dtype=np.uint64), np.array([1, 2], dtype=np.uint32))
    idx.finalize()
    bcs, cts, heads, tails = [], [], [], []
    for b in range(n_bc):
        c = rng.choice(n_ct, per, replace=False)
        h = rng.integers(0, 2, per) * 9  # 9 reads on one end, none on the other: always valid at -c 5, r 0.05
        bcs.append(np.full(per, b)), cts.append(c), heads.append(h), tails.append(9 - h)
    bcs, cts, heads, tails = (np.concatenate(x).astype(np.uint32) for x in (bcs, cts, heads, tails))
    idx.imap_add(bcs, cts, heads, tails)
    rank = rng.permutation(n_ct).astype(np.uint32)
    mult = np.full(n_bc, 100, dtype=np.int32)
    monkeypatch.setenv("ARKS_PMAP_INITIAL_SLOTS", str(1 << 21))
    a, b, c = idx.pair_links(mult, 50, 10000, 5, 0.05, rank)
    # numpy restatement: per barcode all unordered pairs, ordered by rank, orientation 2*(!Ahead)+(!Bhead)
    iu, ju = np.triu_indices(per, 1)
    C = cts.reshape(n_bc, per)
    H = (heads.reshape(n_bc, per) > 0)
    ci, cj, hi, hj = C[:, iu].ravel(), C[:, ju].ravel(), H[:, iu].ravel(), H[:, ju].ravel()
    first = rank[ci] < rank[cj]
    ra, rb = np.where(first, rank[ci], rank[cj]).astype(np.int64), np.where(first, rank[cj], rank[ci]).astype(np.int64)
    ah, bh = np.where(first, hi, hj), np.where(first, hj, hi)
    orient = 2 * (~ah).astype(np.int64) + (~bh).astype(np.int64)
    key = (ra << 32) | rb
    uk, inv = np.unique(key, return_inverse=True)
    want = np.zeros((len(uk), 4), dtype=np.uint32)
    np.add.at(want, (inv, orient), 1)
    inv_rank = np.argsort(rank).astype(np.uint32)
    assert len(a) == len(uk) > 3_000_000
    assert np.array_equal(a, inv_rank[uk >> 32]) and np.array_equal(b, inv_rank[uk & 0xFFFFFFFF])
    assert np.array_equal(c, want)
    assert int(c.sum()) == n_bc * per * (per - 1) // 2
