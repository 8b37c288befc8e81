"""N>1 path on CPU: two gloo ranks each hold the pair links of their barcode shard; the merge
(all-gather keys + all-reduce counters, arcs_b200/merge.py) must reproduce the single-rank link map."""
import os
import socket

import numpy as np
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


def _worker(rank_id, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank_id, world_size=world)
    from arcs_b200.merge import merge_pmap
    rows, mult, lexrank = _workload()
    mine = rows[rows[:, 0] % world == rank_id]  # barcode-sharded
    a, b, c = _links(mine, mult, lexrank)
    ma, mb, mc = merge_pmap(a, b, c, "cpu")
    if rank_id == 0:
        np.savez(out, a=ma, b=mb, c=mc)
    dist.destroy_process_group()


def test_two_rank_merge_equals_single_rank(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    rows, mult, lexrank = _workload()
    a, b, c = _links(rows, mult, lexrank)
    want = {(int(x), int(y)): tuple(int(v) for v in z) for x, y, z in zip(a, b, c)}
    have = {(int(x), int(y)): tuple(int(v) for v in z) for x, y, z in zip(got["a"], got["b"], got["c"])}
    assert len(want) > 50
    assert have == want
