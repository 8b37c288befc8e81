"""At-scale parity through the command line: arcs_b200/bin/arcs (GPU) against the reference's own code
(oracle/_ref/arcs_ref, all host threads) on the SAME FASTA/FASTQ files, bench-shaped workloads of about a million
read pairs -- byte comparison of _original.gv, _main.tsv and the pair map, and the 16 verbose counters
(bench.parity_sample does the comparison; bench.py runs the same check on a sample of its own workload)."""
import os

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "arcs_ref")

SHAPES = {
    # config, draft bases, contigs, read pairs
    "c2": (10_000_000, 1000, 1_000_000),
    "c3": (5_000_000, 100, 1_000_000),
    "c5": (20_000_000, 200, 300_000),
}


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (needs /root/reference once)")
@pytest.mark.parametrize("name", sorted(SHAPES))
def test_cli_equals_reference_code_at_bench_scale(name):
    import numpy as np
    import torch

    import bench
    genome_len, n_contigs, n_pairs = SHAPES[name]
    cfg = dict(bench.CONFIGS[name], genome=genome_len, contigs=n_contigs, pairs=n_pairs, name=name)
    dev = torch.device("cuda", 0)
    genome, starts, ends = bench.make_draft(torch, dev, genome_len, n_contigs, seed=11)
    bases, barcode, mult = bench.make_reads(torch, dev, genome, cfg, n_pairs, seed=12)
    L = cfg["read_len"]
    par, cpu = bench.parity_sample(np, cfg, genome.cpu().numpy(), starts, ends, bases.view(-1, L).cpu().numpy(),
                                   barcode.cpu().numpy(), mult, os.cpu_count() or 1)
    assert par["status"] == "ok", par
    assert all(par["identical"].values()) and par["counters_differing"] == []
    assert par["pair_links"] > 50 and par["gv_edges"] > 0, par  # the comparison is not vacuous
    assert cpu["value"] > 0
