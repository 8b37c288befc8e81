"""The one exchange step of the multi-GPU path: read pairs are sharded by barcode, so
every rank holds the pair-link counters of ITS barcodes only; the link map of the whole run is the
key-wise sum.  All-gather the ranks' pair keys, form the identical sorted union on every rank, then a
single all-reduce (sum) over the dense 4 x n_keys counter vector.  Integer sums: the result does not
depend on the number of ranks or on arrival order.

torch.distributed is plumbing here (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def merge_pmap(a, b, counts, device):
    """a, b: uint32 contig indices; counts: uint32 [n,4] of this rank -> (a, b, counts) of all ranks,
    sorted by (a, b) as integers (the caller re-orders by name rank)."""
    world = dist.get_world_size()
    keys = torch.from_numpy((a.astype(np.int64) << 32) | b.astype(np.int64)).to(device)
    cnt = torch.from_numpy(counts.astype(np.int64)).reshape(-1, 4).to(device)
    n = torch.tensor([keys.numel()], device=device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    mx = max(1, int(max(s.item() for s in sizes)))
    pad = torch.full((mx,), -1, device=device, dtype=torch.int64)
    pad[:keys.numel()] = keys
    gathered = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad)
    union = torch.unique(torch.cat(gathered))
    union = union[union >= 0]
    dense = torch.zeros((union.numel(), 4), device=device, dtype=torch.int64)
    if keys.numel():
        dense[torch.searchsorted(union, keys)] = cnt
    dist.all_reduce(dense, op=dist.ReduceOp.SUM)
    u = union.cpu().numpy()
    return (u >> 32).astype(np.uint32), (u & 0xFFFFFFFF).astype(np.uint32), dense.cpu().numpy().astype(np.uint32)
