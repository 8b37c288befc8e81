"""Multi-GPU plumbing for one process per GPU.

Read pairs are sharded by barcode, every rank holds the same k-mer index, so every rank's pair-link map covers
ITS barcodes only and the link map of the run is the key-wise sum.  The sum itself is `arks_merge_pmap` in the
C ABI (include/arks_b200.h): NCCL all-gather of the ranks' sorted keys, the same sorted union on every rank, ONE
ncclAllReduce (sum, uint32) over the dense 4 x n_union counter vector -- all on the device.  What is left here is
the rendezvous: the NCCL unique id made by rank 0 travels to the other ranks through torch.distributed
(nccl on the GPU box, gloo in the CPU tests), which is plumbing, not data path.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import api


def shard_of_barcode(barcode_id, world):
    """the rank a barcode's read pairs go to (bench.py; the CLI hashes the barcode text, host/ingest.h)"""
    return np.asarray(barcode_id) % world


def exchange_comm_id(device="cpu", make_id=api.comm_unique_id):
    """rank 0's communicator id, delivered to every rank of the default process group"""
    buf = torch.zeros(128, dtype=torch.uint8)
    if dist.get_rank() == 0:
        buf = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8).clone()
    buf = buf.to(device)
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


def init_comm(idx, device):
    """gives `idx` (ArksIndex) its NCCL communicator: rank / size of the default process group"""
    idx.comm_init_rank(exchange_comm_id(device), dist.get_rank(), dist.get_world_size())


def merge_pmap(idx):
    """collective: afterwards every rank's idx.pmap_rows() / pmap_digest() describe the merged map"""
    idx.merge_pmap()
