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


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (needs /root/reference once)")
@pytest.mark.parametrize("k,j", [(30, 0.5), (60, 0.55)])
def test_cli_equals_reference_code_on_an_adversarial_mid_size_draft(k, j, tmp_path):
    """the same comparison on tools/synth's adversarial draft at a few Mbp: N runs, IUPAC codes, lower case, sequence
    duplicated across contigs, palindromic tracts; reads of ragged length with substitutions and Ns, multi-line FASTA"""
    import subprocess

    import numpy as np

    import bench
    from tools import synth
    rng = np.random.default_rng(1000 + k)
    genome, contigs = synth.make_draft(rng, 3_000_000, 9000, k, n_runs=60, palindromes=24, iupac=40)
    rb, roff, bc = synth.make_reads(rng, genome, n_barcodes=600, pairs_per_barcode=200, mol_len=40000, mols_per_barcode=3,
                                    sub_rate=0.004, n_rate=0.002, len_jitter=30)
    fa, fq = str(tmp_path / "draft.fa"), str(tmp_path / "reads.fq")
    with open(fa, "wb") as f:
        for name, s, e in contigs:
            seq = genome[s:e].tobytes()
            f.write(b">" + name.encode() + b" len=%d\n" % (e - s))
            f.write(b"\n".join(seq[a:a + 70] for a in range(0, len(seq), 70)) + b"\n")
    codes = bench.barcode_text(np, np.arange(int(bc.max()) + 1, dtype=np.int64))
    with open(fq, "wb") as f:
        for i in range(len(bc)):
            tag = b" BX:Z:" + codes[bc[i]].tobytes() + b"\n"
            for m in (0, 1):
                seq = rb[roff[2 * i + m]:roff[2 * i + m + 1]].tobytes()
                f.write(b"@p%d/%d" % (i, m + 1) + tag + seq + b"\n+\n" + b"I" * len(seq) + b"\n")
    common = ["-f", fa, "-k", str(k), "-j", str(j), "-c", "3", "-m", "20-10000", "-e", "3000", "-z", "500", "-r", "0.05"]
    g = subprocess.run([bench.ARCS, "--arks", "-v"] + common + ["-b", str(tmp_path / "gpu"), "-P", "--barcode-counts",
                        str(tmp_path / "gpu_bc.tsv"), fq], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert g.returncode == 0, g.stdout[-2000:]
    r = subprocess.run([REF, "-v"] + common + ["-t", str(os.cpu_count() or 1), "-b", str(tmp_path / "ref"), "--tsv",
                        str(tmp_path / "ref_main.tsv"), "--dump-pmap", str(tmp_path / "ref_pair.tsv"), "--barcode-counts",
                        str(tmp_path / "ref_bc.tsv"), "--dist-gv", str(tmp_path / "ref.dist.gv"), fq],
                       stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=900)
    assert r.returncode == 0
    for a, b in (("gpu_original.gv", "ref_original.gv"), ("gpu_main.tsv", "ref_main.tsv"), ("gpu_pair.tsv", "ref_pair.tsv"),
                 ("gpu_bc.tsv", "ref_bc.tsv"), ("gpu.dist.gv", "ref.dist.gv")):
        assert (tmp_path / a).read_bytes() == (tmp_path / b).read_bytes(), a
    gc, rc = bench.grab_counters(g.stdout), bench.grab_counters(r.stdout)
    for name in bench.EXACT_COUNTERS:
        assert gc[name] == rc[name], (name, gc[name], rc[name])
    assert sum(1 for _ in open(tmp_path / "ref_pair.tsv")) > 200 and b"--" in (tmp_path / "ref_original.gv").read_bytes()
