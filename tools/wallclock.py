#!/usr/bin/env python
"""Wall-clock to _original.gv: arcs_b200/bin/arcs (GPU) vs the reference's own code (oracle/_ref/arcs_ref,
all host threads) on the SAME files, and a byte comparison of the outputs.  SURVEY.md 8(d) metric 2/3.

  python tools/wallclock.py [--genome 10000000] [--pairs 2000000] [--k 60]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=10_000_000)
    ap.add_argument("--pairs", type=int, default=2_000_000)
    ap.add_argument("--keep", default=None)
    ap.add_argument("--gz", action="store_true", help="feed both programs the gzip-compressed FASTQ (single-stream zlib inflate on both sides)")
    args = ap.parse_args()
    threads = os.cpu_count() or 1
    tmp = args.keep or tempfile.mkdtemp(prefix="arks_wall_")
    os.makedirs(tmp, exist_ok=True)
    t0 = time.time()
    fa, fq, mult, windows = bench.write_cpu_sample(np, tmp, args.genome, args.pairs, 7)
    if args.gz:
        subprocess.check_call(["gzip", "-1", "-f", fq])
        fq += ".gz"
    gen_s = time.time() - t0
    common = ["-f", fa, "-k", str(bench.K), "-j", str(bench.J), "-c", "5", "-m", "50-10000", "-e", "30000", "-z", "500", "-r", "0.05"]
    t0 = time.time()
    subprocess.check_call([os.path.join(ROOT, "arcs_b200", "bin", "arcs"), "--arks", "-v"] + common + ["-b", os.path.join(tmp, "gpu"), "-P", fq],
                          stdout=open(os.path.join(tmp, "gpu.log"), "w"), stderr=subprocess.STDOUT)
    gpu_s = time.time() - t0
    t0 = time.time()
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "arcs_ref")] + common + ["-t", str(threads), "-b", os.path.join(tmp, "ref"),
                           "--tsv", os.path.join(tmp, "ref_main.tsv"), "--dump-pmap", os.path.join(tmp, "ref_pair.tsv"),
                           "--timing-json", os.path.join(tmp, "ref.json"), fq], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ref_s = time.time() - t0

    def same(a, b):
        return open(os.path.join(tmp, a), "rb").read() == open(os.path.join(tmp, b), "rb").read()

    log = open(os.path.join(tmp, "gpu.log")).read()
    out = {
        "workload": "%d Mbp draft + %d read pairs (2x150 bp), k=%d, %s FASTQ %.2f GB" % (
            args.genome // 1_000_000, args.pairs, bench.K, "gzip-compressed" if args.gz else "uncompressed", os.path.getsize(fq) / 1e9),
        "read_kmers": windows, "host_threads": threads,
        "gpu_wall_s": gpu_s, "reference_wall_s": ref_s, "speedup_wall": ref_s / gpu_s,
        "reference_phases": json.load(open(os.path.join(tmp, "ref.json"))),
        "gpu_log_tail": [l for l in log.splitlines() if l.startswith(("GPU mapping", "wall-clock"))],
        "identical_original_gv": same("gpu_original.gv", "ref_original.gv"),
        "identical_main_tsv": same("gpu_main.tsv", "ref_main.tsv"),
        "identical_pair_map": same("gpu_pair.tsv", "ref_pair.tsv"),
        "generate_s": gen_s,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
