"""N>1 host logic on CPU, two gloo ranks (no GPU): the communicator id made by rank 0 reaches every rank
(arcs_b200/merge.py -- on the GPU box the same code runs over nccl), barcode sharding is a partition, and the
algebra arks_merge_pmap relies on holds: the pair links of barcode-disjoint shards, summed key by key over the
sorted union of their keys, are the pair links of the whole (checked with the oracle's pairContigs).  The device
side of the merge is tested in tests/test_gpu_merge.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O


def _workload():
    rng = np.random.default_rng(3)
    n_bc, n_ct = 400, 40
    rows = []
    for b in range(n_bc):
        for c in rng.choice(n_ct, size=int(rng.integers(1, 7)), replace=False):
            h, t = int(rng.integers(0, 12)), int(rng.integers(0, 12))
            if h + t:
                rows.append((b, int(c), h, t))
    rows = np.array(sorted(rows), dtype=np.uint32)
    mult = rng.integers(10, 200, n_bc).astype(np.int32)
    rank = rng.permutation(n_ct).astype(np.uint32)
    return rows, mult, rank


def _links(rows, mult, rank):
    return O.pair_contigs(rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3], mult, 20, 150, 3, np.float32(0.05), rank)


def merge_model(keys, counts, world_gather, world_sum):
    """arks_merge_pmap's steps with the collectives passed in: all-gather of the sorted keys -> sorted union ->
    dense counters -> one all-reduce"""
    union = np.unique(np.concatenate(world_gather(keys)))
    dense = np.zeros((len(union), 4), dtype=np.int64)
    dense[np.searchsorted(union, keys)] = counts
    return union, world_sum(dense)


def _worker(rank_id, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank_id, world_size=world)
    from arcs_b200 import merge
    comm_id = merge.exchange_comm_id("cpu")  # ncclGetUniqueId works without a GPU
    rows, mult, lexrank = _workload()
    mine = rows[merge.shard_of_barcode(rows[:, 0], world) == rank_id]
    a, b, c = _links(mine, mult, lexrank)
    keys = (lexrank[a].astype(np.int64) << 32) | lexrank[b].astype(np.int64)
    order = np.argsort(keys)

    def gather(k):
        n = torch.tensor([len(k)])
        sizes = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sizes, n)
        pad = torch.full((int(max(s.item() for s in sizes)),), -1, dtype=torch.int64)
        pad[:len(k)] = torch.from_numpy(k)
        got = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(got, pad)
        return [g.numpy()[:int(s.item())] for g, s in zip(got, sizes)]

    def total(dense):
        t = torch.from_numpy(dense)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()

    union, dense = merge_model(keys[order], c[order].astype(np.int64), gather, total)
    ids = [None] * world
    dist.all_gather_object(ids, comm_id)
    if rank_id == 0:
        np.savez(out, union=union, dense=dense, same_id=np.array([i == ids[0] for i in ids]), id_len=len(comm_id),
                 shard_rows=len(mine))
    dist.destroy_process_group()


def test_two_rank_merge_equals_single_rank(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    assert got["same_id"].all() and int(got["id_len"]) == 128
    rows, mult, lexrank = _workload()
    assert 0 < int(got["shard_rows"]) < len(rows)
    a, b, c = _links(rows, mult, lexrank)
    want = {(int(lexrank[x]) << 32) | int(lexrank[y]): tuple(int(v) for v in z) for x, y, z in zip(a, b, c)}
    have = {int(k): tuple(int(v) for v in z) for k, z in zip(got["union"], got["dense"])}
    assert len(want) > 50
    assert have == want
    assert list(got["union"]) == sorted(want)  # the union comes out in std::map<pair<string,string>> order
