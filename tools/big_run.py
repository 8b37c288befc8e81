#!/usr/bin/env python
"""Wall-clock of the command line at scale (BASELINE.json configs[3] shape: human-scale draft, barcode-grouped linked
reads, k=60): writes a synthetic draft (FASTA) and read set (interleaved FASTQ, plain or bgzf) to tmpfs with bench.py's
generators, runs `arcs --arks --gpus N` on them (process start -> exit, phases from its verbose log), optionally runs
the reference's own code (oracle/_ref/arcs_ref, all host threads) on a SAMPLE of the same reads against the same draft
and extrapolates its mapping phase to the full read set.

  python tools/big_run.py --genome 3000000000 --contigs 300000 --pairs 50000000 --gpus 8 [--ref-pairs 500000] [--bgzf]
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--genome", type=int, default=3_000_000_000)
    ap.add_argument("--contigs", type=int, default=300_000)
    ap.add_argument("--pairs", type=int, default=50_000_000)
    ap.add_argument("--gpus", default="1", help="comma-separated list of GPU counts to run the command line with")
    ap.add_argument("--ref-pairs", type=int, default=0, help="run the reference's code on this many of the pairs (0: skip)")
    ap.add_argument("--chunk", type=int, default=2_000_000)
    ap.add_argument("--gzip", action="store_true", help="also run on the gzip (pigz-less: python zlib level 1, one member per chunk) copy")
    ap.add_argument("--keep", default=None)
    args = ap.parse_args()
    cfg = dict(bench.CONFIGS[args.config], genome=args.genome, contigs=args.contigs, pairs=args.pairs, name=args.config)
    L, k = cfg["read_len"], cfg["k"]
    dev = torch.device("cuda", 0)
    tmp = args.keep or tempfile.mkdtemp(prefix="arks_big_", dir=bench.tmp_root())
    os.makedirs(tmp, exist_ok=True)
    t0 = time.time()
    genome, starts, ends = bench.make_draft(torch, dev, cfg["genome"], cfg["contigs"], seed=1)
    genome_np = genome.cpu().numpy()
    fa = os.path.join(tmp, "draft.fa")
    bench.write_draft_fasta(fa, genome_np, starts, ends)
    del genome_np
    fq = os.path.join(tmp, "reads.fq")
    mults = []
    bc_base = 0
    ref_reads, ref_bc = None, None
    gz = open(os.path.join(tmp, "reads.fq.gz"), "wb") if args.gzip else None
    with open(fq, "wb") as f:
        for c0 in range(0, args.pairs, args.chunk):
            n = min(args.chunk, args.pairs - c0)
            bases, barcode, mult = bench.make_reads(torch, dev, genome, cfg, n, seed=1000 + c0 // args.chunk)
            reads_np = bases.view(-1, L).cpu().numpy()
            bc_np = barcode.cpu().numpy().astype(np.int64) + bc_base
            bc_base += len(mult)
            mults.append(mult)
            path = os.path.join(tmp, "chunk.fq")
            bench.write_fastq(np, path, reads_np, bc_np, L)
            with open(path, "rb") as c:
                data = c.read()
            f.write(data)
            if gz is not None:
                import zlib
                co = zlib.compressobj(1, zlib.DEFLATED, 31)
                gz.write(co.compress(data) + co.flush())
            if c0 == 0 and args.ref_pairs:
                m = min(args.ref_pairs, n)
                ref_reads, ref_bc = reads_np[:2 * m].copy(), bc_np[:m].copy()
            os.remove(path)
            del bases, barcode, reads_np, data
    if gz is not None:
        gz.close()
    del genome
    torch.cuda.empty_cache()
    gen_s = time.time() - t0
    fq_gb = os.path.getsize(fq) / 1e9
    windows = 2 * args.pairs * (L - k + 1)
    out = {"workload": "%s shape: %d Mbp draft (%d contigs) + %d read pairs of 2x%d bp, k=%d, j=%.2f; FASTA %.2f GB + FASTQ %.2f GB on tmpfs"
                       % (args.config, cfg["genome"] // 1_000_000, cfg["contigs"], args.pairs, L, k, cfg["j"], os.path.getsize(fa) / 1e9, fq_gb),
           "read_kmers": windows, "host_threads": os.cpu_count(), "generate_s": gen_s, "runs": []}
    common = bench.common_cli_args(cfg, fa)
    inputs = [("plain", fq)] + ([("gzip", fq + ".gz")] if args.gzip else [])
    gv = {}
    for kind, path in inputs:
        for n_gpus in [int(x) for x in args.gpus.split(",")]:
            base = os.path.join(tmp, "gpu%d_%s" % (n_gpus, kind))
            t0 = time.perf_counter()
            p = subprocess.run([bench.ARCS, "--arks", "-v", "--gpus", str(n_gpus)] + common + ["-b", base, "-P", path], stdout=subprocess.PIPE,
                               stderr=subprocess.STDOUT, text=True)
            wall = time.perf_counter() - t0
            phases = [ln for ln in p.stdout.splitlines() if ln.startswith(("GPU mapping", "wall-clock", "Stored read pairs", "Number Kmers Recorded"))]
            run = {"gpus": n_gpus, "input": kind, "rc": p.returncode, "wall_s": wall, "kmers_per_s_wall": windows / wall, "phases": phases}
            if p.returncode == 0:
                gv[(n_gpus, kind)] = open(base + "_original.gv", "rb").read() + open(base + "_pair.tsv", "rb").read()
                run["gv_edges"] = sum(1 for ln in open(base + "_original.gv") if "--" in ln)
                run["pair_links"] = sum(1 for _ in open(base + "_pair.tsv"))
            else:
                run["tail"] = p.stdout[-600:]
            out["runs"].append(run)
            print(json.dumps(run), file=sys.stderr, flush=True)
    out["outputs_identical_across_runs"] = len(set(gv.values())) <= 1
    if args.ref_pairs and ref_reads is not None and os.path.exists(bench.REF):
        rfq, mc = os.path.join(tmp, "ref.fq"), os.path.join(tmp, "mult.csv")
        bench.write_fastq(np, rfq, ref_reads, ref_bc, L)
        bench.write_mult_csv(np, mc, np.unique(ref_bc), np.concatenate(mults))
        threads = os.cpu_count() or 1
        t0 = time.perf_counter()
        r = subprocess.run([bench.REF] + common + ["-u", mc, "-t", str(threads), "-b", os.path.join(tmp, "ref"), "--timing-json",
                            os.path.join(tmp, "ref.json"), rfq], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        ref_wall = time.perf_counter() - t0
        if r.returncode == 0:
            tj = json.load(open(os.path.join(tmp, "ref.json")))
            rate = 2 * len(ref_bc) * (L - k + 1) / tj["t_map_s"]
            out["reference"] = {"threads": threads, "sample_pairs": int(len(ref_bc)), "wall_s": ref_wall,
                                "phases": {x: tj[x] for x in ("t_multiplicity_s", "t_index_s", "t_map_s", "t_pair_s", "t_graph_s")},
                                "map_kmers_per_s": rate,
                                "extrapolated_wall_s_full_read_set": tj["t_index_s"] + windows / rate + tj["t_pair_s"],
                                "note": "the reference's own hot-path code (oracle/_ref); index build measured on the full draft, mapping "
                                        "phase measured on the sample and scaled linearly to the full read set (readBarcodes' extra pass over "
                                        "the reads not included: -u given)"}
    print(json.dumps(out))
    if not args.keep:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
