"""arks-long without the pipe: `arcs --arks --cut L --cut_min M <long reads>` cuts the long reads
into pseudo-linked read pairs while it reads them (arcs_b200/host/long_cut.h) and must give exactly what the
reference's code gave on the output of long-to-linked-pe for the same reads (fixtures of tools/make_fixtures.py
cut, made with oracle/_ref), and what our own CLI gives when that tool's output is piped into it
(bin/arcs-make:300-312)."""
import glob
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "cli_cases", "cut_k20")
ARCS = os.path.join(ROOT, "arcs_b200", "bin", "arcs")
LTLPE = os.path.join(ROOT, "arcs_b200", "bin", "long-to-linked-pe")
CASES = sorted(glob.glob(os.path.join(GOLD, "expected_*_args.json")))


def read(path):
    with open(path, "rb") as f:
        return f.read()


@pytest.mark.parametrize("argsfile", CASES, ids=[os.path.basename(c) for c in CASES])
@pytest.mark.parametrize("mode", ["cut", "cut-two-pass", "cut-sequential", "cut-2gpu-ids", "pipe"])
def test_cut_long_reads_on_the_fly(argsfile, mode, tmp_path):
    spec = json.load(open(argsfile))
    exp = argsfile[:-len("_args.json")]
    L, M = spec["cut"]
    reads = os.path.join(GOLD, spec["reads"])
    args = ["--arks", "-f", os.path.join(GOLD, "draft.fa"), "-b", str(tmp_path / "o"), "--barcode-counts", str(tmp_path / "bc.tsv"), "-P"] + spec["args"]
    if spec["multfile"]:
        args += ["-u", os.path.join(GOLD, spec["multfile"])]
    env = dict(os.environ)
    if mode == "pipe":
        cmd = "%s -l %d -m %d %s 2>/dev/null | %s %s /dev/stdin" % (LTLPE, L, M, reads, ARCS, " ".join(args))
        p = subprocess.run(cmd, shell=True, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, env=env)
    else:
        args += ["--cut", str(L), "--cut_min", str(M)]
        if mode == "cut-two-pass":
            args.append("--two-pass")
        if mode == "cut-sequential":
            env["ARKS_PARSE_THREADS"] = "0"
        if mode == "cut-2gpu-ids":
            import torch
            env["ARKS_GPUS"] = "2"
            if torch.cuda.device_count() < 2:
                env["ARKS_GPUS_SAME_DEVICE"] = "1"
        p = subprocess.run([ARCS] + args + [reads], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert read(tmp_path / "o_original.gv") == read(exp + "_original.gv")
    assert read(tmp_path / "o_main.tsv") == read(exp + "_main.tsv")
    assert read(tmp_path / "o_pair.tsv") == read(exp + "_pmap.txt")
    assert read(tmp_path / "bc.tsv") == read(exp + "_bc.tsv")


def test_long_read_shaped_pairs_match_oracle():
    """BASELINE.json's arks-long configuration in small, through the C ABI: pseudo read pairs of 250 bases cut from
    long reads with 8 % substitutions (so nearly every 20-mer window sits next to an error), k=20, j=0.05, barcode
    = long read; per-pair contig ends and all counters against the CPU restatement"""
    import numpy as np
    from test_gpu_parity import _check_index
    from tools import synth
    k, j, L = 20, 0.05, 250
    rng = np.random.default_rng(4242)
    genome, contigs = synth.make_draft(rng, 200000, 12000, k, n_runs=6, palindromes=3)
    bases, end_off, conreci, _ = synth.contig_end_arrays(genome, contigs, k, min_size=500, end_length=30000)
    idx, km = _check_index(k, bases, end_off, conreci)
    reads, bc = [], []
    for i in range(70):
        n = int(np.clip(rng.lognormal(np.log(8000), 0.6), 2000, 40000))
        a = int(rng.integers(0, len(genome) - n))
        r = genome[a:a + n].copy()
        sub = rng.random(n) < 0.08
        r[sub] = synth.ACGT[rng.integers(0, 4, int(sub.sum()))]
        if i % 7 == 0:
            r[rng.integers(0, n, 3)] = ord("N")
        if i % 2:
            r = synth.revcomp(r)
        step = 2 * L
        for p in range(0, n - step + 1, step):  # the cutting rule of long-to-linked-pe
            reads += [r[p:p + L], synth.revcomp(r[p + L:p + step])]
            bc.append(i)
        rem = n % step
        if rem:
            m = min(L, rem)
            reads += [r[n - rem:n - rem + m], synth.revcomp(r[n - m:n])]
            bc.append(i)
    rb = np.concatenate(reads).astype(np.uint8)
    roff = np.zeros(len(reads) + 1, dtype=np.uint32)
    roff[1:] = np.cumsum([len(x) for x in reads])
    got = idx.map_pairs(rb, roff, np.asarray(bc, dtype=np.uint32), j)
    want, st = km.map_pairs(rb, roff, j)
    assert len(bc) > 1000 and (want != 0).sum() > 200
    assert np.array_equal(got, want), np.nonzero(got != want)[0][:10]
    assert idx.map_stats().as_dict() == st.as_dict()
