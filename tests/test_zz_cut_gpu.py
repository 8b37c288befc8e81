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
