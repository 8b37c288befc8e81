#!/usr/bin/env python
"""bench.py -- ARKS hot path throughput on B200: read k-mers/s through the read->contig
lookup, on BASELINE.json configs[1]: synthetic 50 Mbp draft (5k contigs) +
50 M interleaved linked reads (25 M pairs of 2x150 bp), k=60, j=0.55.

A "step" is one pass of the lookup+vote kernel over the whole read set (in batches of
< 4 Gbases because read offsets are 32-bit).  `value` times the kernel with reads resident
in HBM; `e2e` times the same pass through the C ABI with pinned HOST buffers (H2D copies
inside the timed region).  One process per GPU; at N>1 every rank maps its own shard of
read pairs (barcode-sharded, table replicated: weak scaling) and the sparse pair-link maps
are merged once with NCCL after the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--pairs P] ...
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, J = 60, 0.55
READ_LEN = 150
BYTES_PER_KMER = 32.0 + READ_LEN / (READ_LEN - K + 1) + 4.0 / (2 * (READ_LEN - K + 1))  # 33.67 B


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genome", type=int, default=50_000_000)
    ap.add_argument("--contigs", type=int, default=5000)
    ap.add_argument("--pairs", type=int, default=25_000_000)
    ap.add_argument("--batch-pairs", type=int, default=1_562_500)
    ap.add_argument("--pairs-per-barcode", type=int, default=250)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-genome", type=int, default=5_000_000)
    ap.add_argument("--cpu-pairs-per-thread", type=int, default=20_000)
    return ap.parse_args()


# ------------------------------------------------------------------ synthetic workload (torch, on device)

def make_draft_gpu(torch, dev, genome_len, n_contigs, seed):
    """i.i.d. ACGT genome cut into n_contigs log-normal contigs; 1 % of sequence duplicated
    across contigs (forces value-0 keys), a few N runs.  -> (genome uint8 cuda, starts, ends)"""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    genome = lut[torch.randint(0, 4, (genome_len,), generator=g, device=dev)]
    cg = torch.Generator()
    cg.manual_seed(seed + 1)
    w = torch.exp(torch.randn(n_contigs, generator=cg) * 0.5)
    lens = torch.clamp((w / w.sum() * genome_len).long(), min=1000)
    ends = torch.cumsum(lens, 0)
    ends[-1] = genome_len
    starts = torch.cat([torch.zeros(1, dtype=torch.long), ends[:-1]])
    ends = torch.clamp(ends, max=genome_len)
    n_dup = genome_len // 100 // 200
    src = torch.randint(0, genome_len - 200, (n_dup,), generator=cg)
    dst = torch.randint(0, genome_len - 200, (n_dup,), generator=cg)
    ar = torch.arange(200)
    genome[(dst[:, None] + ar).flatten().to(dev)] = genome[(src[:, None] + ar).flatten().to(dev)]
    for p in torch.randint(0, genome_len - 300, (n_contigs // 200 + 1,), generator=cg).tolist():
        genome[p:p + 1 + (p % 180)] = ord("N")
    return genome, starts, ends


def contig_ends(starts, ends, min_size=500, end_length=30000):
    """getContigKmers' end rule (Arcs.cpp:1056-1091) -> list of (start, stop, conreci) genome intervals"""
    out, i = [], 0
    for s, e in zip(starts.tolist(), ends.tolist()):
        L = e - s
        if L < min_size:
            continue
        cut = end_length
        if cut == 0 or L <= 2 * cut:
            cut = L // 2
        out.append((s, s + cut, 2 * i + 1))
        out.append((e - cut, e, 2 * i + 2))
        i += 1
    return out


def make_reads_gpu(torch, dev, genome, n_pairs, pairs_per_barcode, seed, mol_len=50000, mols=10, insert=350,
                   sub_rate=0.002, n_rate=0.001, chunk=1_000_000):
    """molecule-linked read pairs grouped by barcode -> (bases uint8 [n_pairs*2*READ_LEN] cuda, barcode int32)"""
    G = genome.numel()
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    comp = torch.zeros(256, dtype=torch.uint8, device=dev)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    bases = torch.empty(n_pairs * 2 * READ_LEN, dtype=torch.uint8, device=dev)
    n_bc = (n_pairs + pairs_per_barcode - 1) // pairs_per_barcode
    mol_start = torch.randint(0, G - mol_len, (n_bc, mols), generator=g, device=dev)
    ar = torch.arange(READ_LEN, device=dev)
    for c0 in range(0, n_pairs, chunk):
        n = min(chunk, n_pairs - c0)
        pid = torch.arange(c0, c0 + n, device=dev)
        bc = pid // pairs_per_barcode
        which = torch.randint(0, mols, (n,), generator=g, device=dev)
        ms = mol_start[bc, which]
        p1 = ms + torch.randint(0, mol_len - insert - READ_LEN, (n,), generator=g, device=dev)
        p1 = torch.clamp(p1, max=G - insert - READ_LEN - 1)
        r1 = genome[(p1[:, None] + ar)]
        r2 = comp[genome[(p1[:, None] + insert + ar)].long()].flip(1)
        flip = torch.rand(n, generator=g, device=dev) < 0.5
        a = torch.where(flip[:, None], r2, r1)
        b = torch.where(flip[:, None], r1, r2)
        blk = torch.cat([a, b], dim=1).reshape(-1)
        sub = torch.rand(blk.numel(), generator=g, device=dev) < sub_rate
        blk = torch.where(sub, lut[torch.randint(0, 4, (blk.numel(),), generator=g, device=dev)], blk)
        nn = torch.rand(blk.numel(), generator=g, device=dev) < n_rate
        blk = torch.where(nn, torch.full_like(blk, ord("N")), blk)
        bases[c0 * 2 * READ_LEN:(c0 + n) * 2 * READ_LEN] = blk
    barcode = (torch.arange(n_pairs, device=dev) // pairs_per_barcode).to(torch.int32)
    return bases, barcode


# ------------------------------------------------------------------ clocks

class ClockSampler:
    """samples nvidia-smi during the timed region (B200_PROFILING.md clocks line)"""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].startswith("Active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------ CPU baseline (the reference's own code)

def write_cpu_sample(np, tmp, genome_len, n_pairs, seed):
    """small workload of the same shape, written as FASTA + uncompressed interleaved FASTQ + multiplicity CSV"""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    genome = acgt[rng.integers(0, 4, genome_len)]
    n_contigs = max(2, genome_len // 10000)
    cuts = np.sort(rng.choice(np.arange(1000, genome_len - 1000), n_contigs - 1, replace=False))
    bounds = np.concatenate([[0], cuts, [genome_len]])
    fa = os.path.join(tmp, "draft.fa")
    with open(fa, "wb") as f:
        for i in range(n_contigs):
            f.write(b">%d\n" % (i + 1))
            f.write(genome[bounds[i]:bounds[i + 1]].tobytes())
            f.write(b"\n")
    ppb, mols, mol_len, insert = 250, 10, 50000, 350
    n_bc = (n_pairs + ppb - 1) // ppb
    bc = np.arange(n_pairs) // ppb
    mol_start = rng.integers(0, genome_len - mol_len, (n_bc, mols))
    p1 = mol_start[bc, rng.integers(0, mols, n_pairs)] + rng.integers(0, mol_len - insert - READ_LEN, n_pairs)
    p1 = np.minimum(p1, genome_len - insert - READ_LEN - 1)
    ar = np.arange(READ_LEN)
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    r1 = genome[p1[:, None] + ar]
    r2 = comp[genome[p1[:, None] + insert + ar]][:, ::-1]
    reads = np.stack([r1, r2], axis=1).reshape(-1, READ_LEN).copy()
    sub = rng.random(reads.shape) < 0.002
    reads[sub] = acgt[rng.integers(0, 4, int(sub.sum()))]
    reads[rng.random(reads.shape) < 0.001] = ord("N")
    # fixed-width records so the file can be assembled with numpy
    n_reads = reads.shape[0]
    head = np.frombuffer(b"@r000000000 BX:Z:AAAAAAAAAAAAAAAA-1\n", dtype=np.uint8)
    rec = np.empty((n_reads, len(head) + READ_LEN + 3 + READ_LEN + 1), dtype=np.uint8)
    rec[:, :len(head)] = head
    pid = np.repeat(np.arange(n_pairs), 2)
    for d in range(9):
        rec[:, 2 + d] = ord("0") + (pid // 10 ** (8 - d)) % 10
    bcr = np.repeat(bc, 2)
    for d in range(16):
        rec[:, 17 + d] = acgt[(bcr >> (2 * (15 - d))) & 3]
    o = len(head)
    rec[:, o:o + READ_LEN] = reads
    rec[:, o + READ_LEN:o + READ_LEN + 3] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    rec[:, o + READ_LEN + 3:o + 2 * READ_LEN + 3] = ord("I")
    rec[:, -1] = ord("\n")
    fq = os.path.join(tmp, "reads.fq")
    rec.tofile(fq)
    mult = os.path.join(tmp, "mult.csv")
    with open(mult, "w") as f:
        for b in range(n_bc):
            code = "".join("ACGT"[(b >> (2 * (15 - d))) & 3] for d in range(16))
            f.write("%s-1,%d\n" % (code, 2 * min(ppb, n_pairs - b * ppb)))
    windows = n_reads * (READ_LEN - K + 1)
    return fa, fq, mult, windows


def run_reference_once(fa, fq, mult, tmp, threads):
    ref = os.path.join(ROOT, "oracle", "_ref", "arcs_ref")
    tj = os.path.join(tmp, "timing.json")
    subprocess.check_call([ref, "-f", fa, "-k", str(K), "-j", str(J), "-c", "5", "-m", "50-10000", "-t", str(threads), "-u", mult,
                           "-b", os.path.join(tmp, "out"), "--timing-json", tj, fq],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t = json.load(open(tj))
    return (t["read_kmers_valid"] + t["read_kmers_invalid"]), t["t_map_s"]


def cpu_baseline(np, args, threads):
    ref = os.path.join(ROOT, "oracle", "_ref", "arcs_ref")
    if not os.path.exists(ref):
        return {"value": None, "unit": "k-mers/s", "cores": threads, "kind": "reference", "sample": "oracle/_ref/arcs_ref not built"}
    n_pairs = args.cpu_pairs_per_thread * threads
    with tempfile.TemporaryDirectory() as tmp:
        fa, fq, mult, _ = write_cpu_sample(np, tmp, args.cpu_genome, n_pairs, 99)
        kmers, secs = run_reference_once(fa, fq, mult, tmp, threads)
    return {"value": kmers / secs, "unit": "k-mers/s", "cores": threads, "kind": "reference",
            "sample": "reference's own chromiumRead/bestContig (oracle/_ref, std::unordered_map for sparsehash), "
                      "%d Mbp draft + %d read pairs of the same generator, k=%d, -t %d, uncompressed FASTQ, mapping phase only"
                      % (args.cpu_genome // 1_000_000, n_pairs, K, threads)}


def main_reference(args):
    """the reference's own CPU implementation of the path, all host threads, bounded sample per step"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    threads = os.cpu_count() or 1
    n_pairs = args.cpu_pairs_per_thread * threads
    with tempfile.TemporaryDirectory() as tmp:
        fa, fq, mult, _ = write_cpu_sample(np, tmp, args.cpu_genome, n_pairs, 99)
        for _ in range(args.warmup):
            run_reference_once(fa, fq, mult, tmp, threads)
        tot_k, tot_s = 0, 0.0
        for _ in range(args.steps):
            kmers, secs = run_reference_once(fa, fq, mult, tmp, threads)
            tot_k += kmers
            tot_s += secs
    v = tot_k / tot_s
    sample = ("%d Mbp draft + %d read pairs per step (same generator as the GPU arm's workload), -t %d, uncompressed "
              "FASTQ, mapping phase (readChroms) only" % (args.cpu_genome // 1_000_000, n_pairs, threads))
    print(json.dumps({
        "impl": "reference", "metric": "read k-mers/s (read->contig lookup)", "value": v, "unit": "k-mers/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * tot_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64 (2-bit packed k-mers)",
        "data": "synthetic",
        "config": {"workload": "synthetic 50 Mbp draft (5k contigs) + 50M interleaved linked reads, k=60, j=0.55 "
                               "(bounded CPU sample: " + sample + ")"},
        "cpu_baseline": {"value": v, "unit": "k-mers/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------ GPU arm

def main():
    args = parse_args()
    if args.impl == "reference":
        return main_reference(args)
    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    import arcs_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- workload: same draft on every rank (replicated table), rank-specific read shard
    genome, starts, ends = make_draft_gpu(torch, dev, args.genome, args.contigs, seed=1)
    iv = contig_ends(starts, ends)
    end_bases = torch.cat([genome[s:e] for s, e, _ in iv])
    h_end_off = np.zeros(len(iv) + 1, dtype=np.uint64)
    h_end_off[1:] = np.cumsum([e - s for s, e, _ in iv])
    d_end_off = torch.from_numpy(h_end_off.astype(np.int64)).to(dev)
    d_conreci = torch.tensor([c for _, _, c in iv], dtype=torch.int32, device=dev)
    n_contigs = len(iv) // 2

    idx = arcs_b200.ArksIndex(K, int(h_end_off[-1]), device=local)
    # the kernels must run on the stream the CUDA events are recorded on: a real (non-default) torch stream,
    # made current so that the workload generation, the events and the library all use it (a null stream
    # handle would make the library fall back to its own stream, which torch's events do not see)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.current_stream().synchronize()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    idx.set_stream(stream.cuda_stream)
    t0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    idx.add_ends_device(end_bases.data_ptr(), d_end_off.data_ptr(), d_conreci.data_ptr(), h_end_off)
    ist = idx.finalize().as_dict()
    ev1.record()
    torch.cuda.synchronize()
    index_ms = ev0.elapsed_time(ev1)
    del end_bases

    bases, barcode = make_reads_gpu(torch, dev, genome, args.pairs, args.pairs_per_barcode, seed=2 + rank)
    barcode += rank * ((args.pairs + args.pairs_per_barcode - 1) // args.pairs_per_barcode)  # barcode-disjoint shards
    n_pairs = args.pairs
    bp = min(args.batch_pairs, n_pairs)
    assert bp * 2 * READ_LEN < 2 ** 32
    n_batches = (n_pairs + bp - 1) // bp
    # per-batch offsets are relative to the batch's first base
    off = (torch.arange(2 * bp + 1, device=dev, dtype=torch.int64) * READ_LEN).to(torch.int32)
    torch.cuda.synchronize()

    def step_device():
        for b in range(n_batches):
            a = b * bp
            n = min(bp, n_pairs - a)
            idx.map_pairs_device(bases.data_ptr() + a * 2 * READ_LEN, off.data_ptr(), barcode.data_ptr() + 4 * a, n,
                                 n * 2 * READ_LEN, J, None)

    # one counted pass to learn the work per step
    idx.map_stats_reset()
    step_device()
    st = idx.map_stats().as_dict()
    kmers_per_step = st["kmers_valid"] + st["kmers_invalid"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(2, args.warmup - 1)):  # never fewer than three untimed passes (the counted one included)
        step_device()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = idx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1000
    ms = e0.elapsed_time(e1)
    # the events bracket the kernels on their own stream; the host clock around the same region (which ends
    # with a synchronize) can only be longer -- if it is much longer the events missed the work
    if ms < 0.5 * wall_ms - 1.0:
        raise SystemExit("bench.py: CUDA events (%.3f ms) do not cover the timed region (%.3f ms wall)" % (ms, wall_ms))
    gpu_launches = idx.launches - launches0
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    k_all = torch.tensor([float(kmers_per_step)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(k_all, op=dist.ReduceOp.SUM)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = float(k_all.item()) / (ms_per_step / 1000.0)

    # roofline of the lookup kernels: one batch = map_groups_kernel (lane-per-read fast path) + map_slow_kernel
    # (the deferred mates); algorithmic bytes of the batch / measured duration of the pair of launches (this rank)
    launch_ms = ms / (args.steps * n_batches)
    achieved = BYTES_PER_KMER * (kmers_per_step / n_batches) / (launch_ms / 1000.0) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    traffic = None
    tp = os.path.join(ROOT, "profiles", "map_kernel_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")

    # ---- end to end: pinned host buffers through arks_map_pairs (H2D inside the timed region)
    e2e = None
    if not args.no_e2e:
        L = arcs_b200.load_library()
        nbytes = n_pairs * 2 * READ_LEN
        hp = C.c_void_p()
        assert L.arks_host_alloc(C.byref(hp), nbytes) == 0
        hb = C.c_void_p()
        assert L.arks_host_alloc(C.byref(hb), 4 * n_pairs) == 0
        ho = C.c_void_p()
        assert L.arks_host_alloc(C.byref(ho), 4 * (2 * bp + 1)) == 0
        torch.cuda.synchronize()
        # fill the pinned buffers from the device copy (outside the timed region)
        h_bases = torch.frombuffer((C.c_uint8 * nbytes).from_address(hp.value), dtype=torch.uint8)
        h_bases.copy_(bases)
        h_bc = torch.frombuffer((C.c_uint8 * (4 * n_pairs)).from_address(hb.value), dtype=torch.int32)
        h_bc.copy_(barcode)
        h_off = torch.frombuffer((C.c_uint8 * (4 * (2 * bp + 1))).from_address(ho.value), dtype=torch.int32)
        h_off.copy_(off)
        torch.cuda.synchronize()
        stats_host = arcs_b200.MapStats()

        def step_host():
            for b in range(n_batches):
                a = b * bp
                n = min(bp, n_pairs - a)
                idx.map_pairs_raw(hp.value + a * 2 * READ_LEN, ho.value, hb.value + 4 * a, n, J)
            return idx.map_stats()  # device->host read of the step's result (counters); synchronises

        for _ in range(2):
            step_host()
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            stats_host = step_host()
        e1.record()
        barrier()
        e2e_ms = max(e0.elapsed_time(e1), 0.0)
        wall_ms = (time.perf_counter() - t0) * 1000
        t = torch.tensor([max(e2e_ms, wall_ms)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_step_s = float(t.item()) / args.steps / 1000.0
        e2e = {"value": float(k_all.item()) / e2e_step_s, "unit": "k-mers/s",
               "h2d_bytes_per_step": int(nbytes + 4 * n_pairs + n_batches * 4 * (2 * bp + 1)),
               "d2h_bytes_per_step": C.sizeof(arcs_b200.MapStats) + 8, "ms_per_step": e2e_step_s * 1000}
        del stats_host
        L.arks_host_free(hp)
        L.arks_host_free(hb)
        L.arks_host_free(ho)

    # ---- pair links once, and (N>1) the single NCCL merge of the sparse pair-link map
    n_bc_total = int(barcode.max().item()) + 1
    mult = np.full(n_bc_total, 2 * args.pairs_per_barcode, dtype=np.int32)
    rankv = np.argsort(np.argsort(np.array([str(i + 1) for i in range(n_contigs)]))).astype(np.uint32)
    torch.cuda.synchronize()
    tl0 = time.perf_counter()
    pa, pb, pc = idx.pair_links(mult, 50, 10000, 5, 0.05, rankv)
    links_ms = (time.perf_counter() - tl0) * 1000
    merge_ms, merged_pairs = None, int(len(pa))
    if world > 1:
        tm0 = time.perf_counter()
        merged_pairs = nccl_merge_pmap(torch, dist, dev, pa, pb, pc, n_contigs)
        merge_ms = (time.perf_counter() - tm0) * 1000

    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:
            cpu = cpu_baseline(np, args, os.cpu_count() or 1)
        out = {
            "metric": "read k-mers/s (read->contig lookup)", "value": value, "unit": "k-mers/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64 (2-bit packed k-mers)", "data": "synthetic",
            "config": {
                "workload": "synthetic %d Mbp draft (%d contigs) + %dM interleaved linked reads (%d pairs of 2x%d bp) per GPU, "
                            "k=%d, j=%.2f" % (args.genome // 1_000_000, args.contigs, 2 * n_pairs // 1_000_000, n_pairs,
                                              READ_LEN, K, J),
                "l2": "inputs larger than L2 (reads %.1f GB + table %.1f GB per pass)" % (
                    n_pairs * 2 * READ_LEN / 1e9, ist["recorded"] * 2 * 32 / 1e9),
                "batches_per_step": n_batches, "kmers_per_step_per_gpu": kmers_per_step, "table_keys": ist["recorded"],
                "index_build_ms": index_ms, "pair_links_ms": links_ms, "pmap_merge_ms": merge_ms, "pmap_pairs": merged_pairs,
                "parallelism": "barcode-sharded reads x%d, replicated k-mer table" % world,
            },
            "clocks": clk, "gpu_launches": int(gpu_launches), "e2e": e2e,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "map_groups_kernel<2> (lane-per-read, 95 % of the batch) + map_slow_kernel<2> (deferred mates); one batch", "bytes_per_kmer": BYTES_PER_KMER,
                         "peak_source": peak_src, "launch_ms": launch_ms},
            "cpu_baseline": cpu,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def nccl_merge_pmap(torch, dist, dev, pa, pb, pc, n_contigs):
    """the single NCCL exchange: see arcs_b200/merge.py"""
    from arcs_b200.merge import merge_pmap
    a, b, c = merge_pmap(pa, pb, pc, dev)
    torch.cuda.synchronize()
    return int(len(a))


if __name__ == "__main__":
    main()
