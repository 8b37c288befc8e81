#!/usr/bin/env python
"""bench.py -- ARKS hot path throughput on B200: read k-mers/s through the read->contig lookup.

Workloads (--config; BASELINE.json `configs`, sizes per GPU):
  c2  configs[1]  50 Mbp draft (5k contigs) + 50 M interleaved linked reads (25 M pairs of 2x150 bp), k=60, j=0.55,
                  barcode-grouped file, -m 50-10000                                                  [default]
  c3  configs[2]  5 Mbp draft (100 contigs) + 10 M stLFR-style reads (5 M pairs of 2x150 bp), k=40, barcodes
                  shuffled over the file, -m 2-10000
  c5  configs[4]  arks-long: 1 Gbp draft (10k contigs) + ONT-like long reads (8 % error: substitutions, insertions,
                  deletions) cut into 2x250 bp pairs by the long-to-linked-pe rule, k=20, j=0.05, -c 4 -m 8-10000

A "step" is one pass of the lookup+vote kernels over the whole read set (in batches of < 4 Gbases: read offsets
are 32-bit).  `value` times the kernels with the reads resident in HBM; `e2e` times the same pass through the C ABI
(arks_map_pairs) from pinned HOST buffers, copies inside the timed region.  `job` times what follows the pass in a
run of the pipeline: one pass + pair links (pairContigs on the device, ordered rows) + -- at N > 1 -- the NCCL
merge of the sparse pair-link maps.  One process per GPU; at N>1 every rank maps its own barcode-disjoint shard
(table replicated: weak scaling).  Outside the timed regions, at N=1: the parity sample (the command line on a
sample of this very workload against the reference's own code, oracle/_ref/arcs_ref, byte for byte) which also
yields the CPU baseline, and at every N the invariance workload (a fixed set of pairs split by barcode over the
ranks; the digest of the merged pair-link map must not depend on N).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c5] [--impl reference] [--strong] ...
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

CONFIGS = {
    "c2": dict(
        title="synthetic 50 Mbp draft (5k contigs) + 50M interleaved linked reads, k=60, j=0.55",
        genome=50_000_000, contigs=5000, pairs=25_000_000, read_len=150, k=60, j=0.55, kind="linked", ppb=250, mols=10,
        mol_len=50000, shuffled=False, min_mult=50, max_mult=10000, min_reads=5, batch_pairs=3_125_000, cpu_genome=50_000_000),
    "c3": dict(
        title="E. coli-scale 5 Mbp draft + 10M stLFR-style barcoded reads, k=40, -m 2-10000",
        genome=5_000_000, contigs=100, pairs=5_000_000, read_len=150, k=40, j=0.55, kind="linked", ppb=10, mols=1,
        mol_len=50000, shuffled=True, min_mult=2, max_mult=10000, min_reads=5, batch_pairs=1_250_000, cpu_genome=5_000_000),
    "c5": dict(
        title="arks-long: synthetic ONT long reads over a 1 Gbp draft via long-to-linked-pe segmentation, k=20, j=0.05",
        genome=1_000_000_000, contigs=10000, pairs=10_000_000, read_len=250, k=20, j=0.05, kind="long", long_mean=12000,
        long_sigma=0.7, long_min=2000, err_sub=0.04, err_ins=0.02, err_del=0.02, shuffled=False, min_mult=8, max_mult=10000,
        min_reads=4, batch_pairs=1_000_000, cpu_genome=50_000_000),
}
END_LENGTH, MIN_SIZE, ERROR_PERCENT = 30000, 500, 0.05


def bytes_per_kmer(L, k):
    """one 32-byte slot + the read text (1 B per base) + 4 B of output per pair"""
    return 32.0 + L / (L - k + 1) + 4.0 / (2 * (L - k + 1))


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--genome", type=int, default=None)
    ap.add_argument("--contigs", type=int, default=None)
    ap.add_argument("--pairs", type=int, default=None)
    ap.add_argument("--batch-pairs", type=int, default=None)
    ap.add_argument("--strong", action="store_true", help="the configured pairs are split over the ranks by barcode (fixed total work)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the parity sample / CPU baseline")
    ap.add_argument("--no-job", action="store_true")
    ap.add_argument("--no-cli", action="store_true", help="skip the command-line wall-clock leg")
    ap.add_argument("--parity-pairs", type=int, default=320_000)
    ap.add_argument("--cli-pairs", type=int, default=2_000_000)
    ap.add_argument("--invariance-pairs", type=int, default=2_000_000)
    ap.add_argument("--cpu-genome", type=int, default=None)
    ap.add_argument("--cpu-pairs-per-thread", type=int, default=20_000)
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    for key in ("genome", "contigs", "pairs", "batch_pairs", "cpu_genome"):
        if getattr(args, key) is not None:
            cfg[key] = getattr(args, key)
    cfg["name"] = args.config
    return args, cfg


# ------------------------------------------------------------------ synthetic workload (torch; on the device when there is one)

def make_draft(torch, dev, genome_len, n_contigs, seed):
    """i.i.d. ACGT genome cut into n_contigs log-normal contigs; 1 % of the sequence duplicated across contigs
    (forces value-0 keys), a few N runs.  Deterministic: the duplicated segments are read before any is written and
    their destinations do not overlap.  -> (genome uint8, starts, ends)"""
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
    dst = torch.randperm(genome_len // 400 - 1, generator=cg)[:n_dup] * 400  # disjoint 200-base destinations
    ar = torch.arange(200)
    vals = genome[(src[:, None] + ar).flatten().to(dev)].clone()
    genome[(dst[:, None] + ar).flatten().to(dev)] = vals
    for p in torch.randint(0, genome_len - 300, (n_contigs // 200 + 1,), generator=cg).tolist():
        genome[p:p + 1 + (p % 180)] = ord("N")
    return genome, starts, ends


def digest(torch, t):
    """order-sensitive checksum of a byte tensor (for the run manifest)"""
    x = t.to(torch.int64)
    w = (torch.arange(x.numel(), device=t.device, dtype=torch.int64) % 65521) + 1
    return int((x * w).sum().item() & 0x7FFFFFFFFFFFFFFF)


def contig_ends(starts, ends, min_size=MIN_SIZE, end_length=END_LENGTH):
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


def _comp_lut(torch, dev):
    comp = torch.zeros(256, dtype=torch.uint8, device=dev)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    return comp


def make_linked_reads(torch, dev, genome, cfg, n_pairs, seed, insert=350, sub_rate=0.002, n_rate=0.001, chunk=1_000_000):
    """molecule-linked 2xL read pairs -> (bases uint8 [n_pairs*2*L], barcode int32 [n_pairs], mult int32 [n_barcodes]);
    grouped by barcode, or (cfg.shuffled, stLFR-style) in random order"""
    L, ppb, mols, mol_len = cfg["read_len"], cfg["ppb"], cfg["mols"], cfg["mol_len"]
    G = genome.numel()
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    comp = _comp_lut(torch, dev)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    bases = torch.empty(n_pairs * 2 * L, dtype=torch.uint8, device=dev)
    n_bc = (n_pairs + ppb - 1) // ppb
    mol_start = torch.randint(0, G - mol_len, (n_bc, mols), generator=g, device=dev)
    ar = torch.arange(L, device=dev)
    for c0 in range(0, n_pairs, chunk):
        n = min(chunk, n_pairs - c0)
        pid = torch.arange(c0, c0 + n, device=dev)
        bc = pid // ppb
        which = torch.randint(0, mols, (n,), generator=g, device=dev)
        ms = mol_start[bc, which]
        p1 = ms + torch.randint(0, mol_len - insert - L, (n,), generator=g, device=dev)
        p1 = torch.clamp(p1, max=G - insert - L - 1)
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
        bases[c0 * 2 * L:(c0 + n) * 2 * L] = blk
    barcode = (torch.arange(n_pairs, device=dev) // ppb).to(torch.int32)
    if cfg["shuffled"]:
        perm = torch.randperm(n_pairs, generator=g, device=dev)
        bases = bases.view(n_pairs, 2 * L)[perm].reshape(-1)
        barcode = barcode[perm].contiguous()
    mult = (2 * torch.bincount(barcode.long(), minlength=n_bc)).to(torch.int32).cpu().numpy()
    return bases, barcode, mult


def make_long_read_pairs(torch, dev, genome, cfg, n_pairs, seed, chunk=200_000):
    """ONT-like long reads (log-normal lengths, either strand, substitution + insertion + deletion errors) cut by
    long-to-linked-pe's rule (src/long-to-linked-pe.cpp:191-253): for every long read of at least `m` bases, pairs at
    i = 0, 2l, 4l, ...: mate 1 = seq[i, i+l), mate 2 = reverse complement of seq[i+l, i+2l); barcode = the long
    read.  (The ragged remainder pair of each long read is left out: the generator keeps all reads at 2 x l.)"""
    import numpy as np
    L = cfg["read_len"]
    G = genome.numel()
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    rng = np.random.default_rng(seed)
    # long reads until their segments add up to n_pairs
    lens, segs = [], 0
    while segs < n_pairs:
        x = np.clip(rng.lognormal(np.log(cfg["long_mean"]), cfg["long_sigma"], 1 << 16), cfg["long_min"], 200000).astype(np.int64)
        lens.append(x)
        segs += int((x // (2 * L)).sum())
    lens = np.concatenate(lens)
    nseg = lens // (2 * L)
    cum = np.cumsum(nseg)
    n_long = int(np.searchsorted(cum, n_pairs) + 1)
    lens, nseg, cum = lens[:n_long], nseg[:n_long], cum[:n_long]
    first = cum - nseg
    d_first = torch.from_numpy(first).to(dev)
    d_len = torch.from_numpy(lens).to(dev)
    start = torch.from_numpy(rng.integers(0, G - 200001, n_long)).to(dev)
    rev = torch.from_numpy(rng.random(n_long) < 0.5).to(dev)
    comp = _comp_lut(torch, dev)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    bases = torch.empty(n_pairs * 2 * L, dtype=torch.uint8, device=dev)
    barcode = torch.empty(n_pairs, dtype=torch.int32, device=dev)
    d_cum = torch.from_numpy(cum).to(dev)
    for c0 in range(0, n_pairs, chunk):
        n = min(chunk, n_pairs - c0)
        pid = torch.arange(c0, c0 + n, device=dev)
        lr = torch.searchsorted(d_cum, pid, right=True)
        seg = pid - d_first[lr]
        # the 2l bases of the segment in long-read orientation come from this genome window
        lo = torch.where(rev[lr], start[lr] + d_len[lr] - (seg + 1) * 2 * L, start[lr] + seg * 2 * L)
        ins = torch.rand((n, 2 * L), generator=g, device=dev) < cfg["err_ins"]
        dele = torch.rand((n, 2 * L), generator=g, device=dev) < cfg["err_del"]
        consumed = torch.cumsum((~ins).long(), 1) - (~ins).long() + torch.cumsum(dele.long(), 1)
        src = genome[lo[:, None] + torch.clamp(consumed, max=2 * L + 60)]
        rnd = lut[torch.randint(0, 4, (n, 2 * L), generator=g, device=dev)]
        sub = torch.rand((n, 2 * L), generator=g, device=dev) < cfg["err_sub"]
        fwd = torch.where(ins | sub, rnd, src)                      # genome-forward window with errors
        seq = torch.where(rev[lr][:, None], comp[fwd.long()].flip(1), fwd)  # the segment as the long read has it
        m1 = seq[:, :L]
        m2 = comp[seq[:, L:].long()].flip(1)
        bases[c0 * 2 * L:(c0 + n) * 2 * L] = torch.cat([m1, m2], dim=1).reshape(-1)
        barcode[c0:c0 + n] = lr.to(torch.int32)
        del ins, dele, consumed, src, rnd, sub, fwd, seq
    mult = (2 * torch.bincount(barcode.long(), minlength=n_long)).to(torch.int32).cpu().numpy()
    return bases, barcode, mult


def make_reads(torch, dev, genome, cfg, n_pairs, seed):
    if cfg["kind"] == "long":
        return make_long_read_pairs(torch, dev, genome, cfg, n_pairs, seed)
    return make_linked_reads(torch, dev, genome, cfg, n_pairs, seed)


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


# ------------------------------------------------------------------ files for the command line / the reference's code

def barcode_text(np, ids):
    """16-mer + "-1" per barcode id -> uint8 [n, 18]"""
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = np.empty((len(ids), 18), dtype=np.uint8)
    for d in range(16):
        out[:, d] = acgt[(ids >> (2 * (15 - d))) & 3]
    out[:, 16] = ord("-")
    out[:, 17] = ord("1")
    return out


def write_draft_fasta(path, genome_np, starts, ends):
    with open(path, "wb") as f:
        for i, (s, e) in enumerate(zip(starts.tolist(), ends.tolist())):
            f.write(b">%d\n" % (i + 1))
            f.write(genome_np[s:e].tobytes())
            f.write(b"\n")


def write_fastq(np, path, reads, bc, L):
    """interleaved strict FASTQ, fixed-width records assembled with numpy: reads uint8 [2n, L], bc int [n]"""
    n_reads = reads.shape[0]
    head = np.frombuffer(b"@r000000000 BX:Z:AAAAAAAAAAAAAAAA-1\n", dtype=np.uint8)
    rec = np.empty((n_reads, len(head) + L + 3 + L + 1), dtype=np.uint8)
    rec[:, :len(head)] = head
    pid = np.repeat(np.arange(n_reads // 2), 2)
    for d in range(9):
        rec[:, 2 + d] = ord("0") + (pid // 10 ** (8 - d)) % 10
    rec[:, 17:35] = barcode_text(np, np.repeat(np.asarray(bc, dtype=np.int64), 2))
    o = len(head)
    rec[:, o:o + L] = reads
    rec[:, o + L:o + L + 3] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    rec[:, o + L + 3:o + 2 * L + 3] = ord("I")
    rec[:, -1] = ord("\n")
    rec.tofile(path)
    return rec.size


def write_mult_csv(np, path, ids, mult):
    txt = barcode_text(np, np.asarray(ids, dtype=np.int64))
    with open(path, "w") as f:
        for row, b in zip(txt, ids):
            f.write("%s,%d\n" % (row.tobytes().decode(), int(mult[b])))


def tmp_root():
    return "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None


def common_cli_args(cfg, fa):
    return ["-f", fa, "-k", str(cfg["k"]), "-j", str(cfg["j"]), "-c", str(cfg["min_reads"]), "-m",
            "%d-%d" % (cfg["min_mult"], cfg["max_mult"]), "-e", str(END_LENGTH), "-z", str(MIN_SIZE), "-r", str(ERROR_PERCENT)]


REF = os.path.join(ROOT, "oracle", "_ref", "arcs_ref")
ARCS = os.path.join(ROOT, "arcs_b200", "bin", "arcs")

# verbose counters both programs print (Arcs.cpp:1107-1128,1321-1340); the last two are incremented without
# `omp atomic` upstream (Arcs.cpp:1006-1012), so at -t > 1 the reference may lose a few counts
EXACT_COUNTERS = ["Total number of Kmers:", "Number Null Kmers:", "Number Kmers Recorded:", "Number Kmer Collisions:",
                  "Number Times Kmers Removed (since duplicate in different contig):", "Number of unique kmers (only one contig):",
                  "Stored read pairs:", "Skipped invalid read pairs:", "Skipped reads pairs without a good contig:",
                  "Total valid kmers:", "Number invalid kmers:", "Number of kmers found in ContigKmap:",
                  "Number of kmers recorded in Ktrack:", "Number of kmers found in ContigKmap but duplicate:"]
RACY_COUNTERS = ["Number of reads passing jaccard threshold:", "Number of reads failing jaccard threshold:"]


def grab_counters(text):
    out = {}
    for line in text.splitlines():
        for name in EXACT_COUNTERS + RACY_COUNTERS:
            if line.startswith(name):
                out[name] = int(line[len(name):].split()[0])
    return out


def parity_sample(np, cfg, genome_np, starts, ends, reads_np, bc_np, mult, threads):
    """The command line (arcs_b200/bin/arcs, GPU) and the reference's own code (oracle/_ref/arcs_ref, all host threads)
    on the same files: the bench's own draft + a sample of its own reads.  Byte comparison of _original.gv, _main.tsv
    and the pair map, the 16 verbose counters; the reference's mapping phase is the CPU baseline."""
    if not os.path.exists(REF):
        return {"status": "oracle/_ref/arcs_ref not built"}, None
    L, k = cfg["read_len"], cfg["k"]
    with tempfile.TemporaryDirectory(dir=tmp_root()) as tmp:
        fa, fq, mc = os.path.join(tmp, "draft.fa"), os.path.join(tmp, "reads.fq"), os.path.join(tmp, "mult.csv")
        write_draft_fasta(fa, genome_np, starts, ends)
        write_fastq(np, fq, reads_np, bc_np, L)
        write_mult_csv(np, mc, np.unique(bc_np), mult)
        common = common_cli_args(cfg, fa) + ["-u", mc]
        t0 = time.perf_counter()
        g = subprocess.run([ARCS, "--arks", "-v"] + common + ["-b", os.path.join(tmp, "gpu"), "-P", fq], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True)
        gpu_s = time.perf_counter() - t0
        if g.returncode != 0:
            return {"status": "FAILED: arcs exited %d: %s" % (g.returncode, g.stdout[-400:])}, None
        t0 = time.perf_counter()
        r = subprocess.run([REF, "-v"] + common + ["-t", str(threads), "-b", os.path.join(tmp, "ref"), "--tsv", os.path.join(tmp, "ref_main.tsv"),
                            "--dump-pmap", os.path.join(tmp, "ref_pair.tsv"), "--timing-json", os.path.join(tmp, "ref.json"), fq],
                           stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        ref_s = time.perf_counter() - t0
        if r.returncode != 0:
            return {"status": "FAILED: arcs_ref exited %d" % r.returncode}, None

        def same(a, b):
            return open(os.path.join(tmp, a), "rb").read() == open(os.path.join(tmp, b), "rb").read()

        files = {"_original.gv": same("gpu_original.gv", "ref_original.gv"), "_main.tsv": same("gpu_main.tsv", "ref_main.tsv"),
                 "pair map": same("gpu_pair.tsv", "ref_pair.tsv")}
        gc, rc = grab_counters(g.stdout), grab_counters(r.stdout)
        bad = [n for n in EXACT_COUNTERS if gc.get(n) is None or gc.get(n) != rc.get(n)]
        for n in RACY_COUNTERS:
            if gc.get(n) is None or rc.get(n) is None or not (0.98 * gc[n] <= rc[n] <= gc[n]):
                bad.append(n)
        tj = json.load(open(os.path.join(tmp, "ref.json")))
        n_links = sum(1 for _ in open(os.path.join(tmp, "ref_pair.tsv")))
        n_edges = sum(1 for ln in open(os.path.join(tmp, "ref_original.gv")) if "--" in ln)
        ok = all(files.values()) and not bad
        windows = reads_np.shape[0] * (L - k + 1)
        phases = [ln for ln in g.stdout.splitlines() if ln.startswith(("GPU mapping", "wall-clock"))]
        par = {"status": "ok" if ok else "MISMATCH", "pairs": int(len(bc_np)), "draft_mbp": len(genome_np) // 1_000_000,
               "identical": files, "counters_compared": len(EXACT_COUNTERS) + len(RACY_COUNTERS), "counters_differing": bad,
               "pair_links": n_links, "gv_edges": n_edges, "cli_wall_s": gpu_s, "reference_wall_s": ref_s,
               "reference_threads": threads, "wall_ratio": ref_s / gpu_s, "cli_phases": phases,
               "note": "the command line vs the reference's own code (oracle/_ref/arcs_ref) on the same FASTA/FASTQ files; "
                       "the two Jaccard counters are compared within 2 % (not atomic upstream, Arcs.cpp:1006-1012)"}
        cpu = {"value": windows / tj["t_map_s"], "unit": "k-mers/s", "cores": threads, "kind": "reference",
               "sample": "reference's own chromiumRead/bestContig (oracle/_ref: unmodified hot-path code, std::unordered_map for "
                         "sparsehash), the bench's own %d Mbp draft + the first %d read pairs of its read set, k=%d, -t %d, "
                         "uncompressed FASTQ from tmpfs, mapping phase (readChroms: kseq parse + BX + bestContig) only; index "
                         "build %.1f s and pairContigs %.2f s not included" % (len(genome_np) // 1_000_000, len(bc_np), k, threads,
                                                                             tj["t_index_s"], tj["t_pair_s"]),
               "reference_phases": {x: tj[x] for x in ("t_multiplicity_s", "t_index_s", "t_map_s", "t_pair_s", "t_graph_s")}}
        return par, cpu


def cli_wall(np, cfg, genome_np, starts, ends, reads_np, bc_np, mult):
    """process start -> _original.gv closed, through the command line, plain FASTQ on tmpfs"""
    L, k = cfg["read_len"], cfg["k"]
    with tempfile.TemporaryDirectory(dir=tmp_root()) as tmp:
        fa, fq = os.path.join(tmp, "draft.fa"), os.path.join(tmp, "reads.fq")
        write_draft_fasta(fa, genome_np, starts, ends)
        nbytes = write_fastq(np, fq, reads_np, bc_np, L)
        t0 = time.perf_counter()
        g = subprocess.run([ARCS, "--arks", "-v"] + common_cli_args(cfg, fa) + ["-b", os.path.join(tmp, "gpu"), fq],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        wall = time.perf_counter() - t0
        if g.returncode != 0:
            return {"status": "FAILED: %s" % g.stdout[-300:]}
        windows = reads_np.shape[0] * (L - k + 1)
        phases = [ln for ln in g.stdout.splitlines() if ln.startswith(("GPU mapping", "wall-clock"))]
        reads_s = None
        for ln in phases:
            if ln.startswith("GPU mapping:"):
                reads_s = float(ln.split()[2])
        return {"status": "ok", "pairs": int(len(bc_np)), "fastq_gb": nbytes / 1e9, "wall_s": wall, "reads_phase_s": reads_s,
                "ingest_gb_per_s": nbytes / 1e9 / reads_s if reads_s else None,
                "kmers_per_s_reads_phase": windows / reads_s if reads_s else None, "kmers_per_s_wall": windows / wall,
                "phases": phases, "note": "one-pass (multiplicities counted while mapping), FASTQ parse + BX extraction + barcode "
                                          "interning + H2D + kernels + pair links + graph, process start to exit"}


# ------------------------------------------------------------------ the reference arm

def main_reference(args, cfg):
    """the reference's own CPU implementation of the path (oracle/_ref/arcs_ref), all host threads; a step = the
    mapping phase (readChroms) over a bounded sample of the workload; the index is built once (--map-repeats)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch
    threads = os.cpu_count() or 1
    n_pairs = args.cpu_pairs_per_thread * threads
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    cg = dict(cfg)
    cg["genome"] = cfg["cpu_genome"]
    cg["contigs"] = max(2, cfg["contigs"] * cg["genome"] // cfg["genome"])
    genome, starts, ends = make_draft(torch, dev, cg["genome"], cg["contigs"], seed=1)
    bases, barcode, mult = make_reads(torch, dev, genome, cg, n_pairs, seed=2)
    L, k = cfg["read_len"], cfg["k"]
    reads_np = bases.view(-1, L).cpu().numpy()
    bc_np = barcode.cpu().numpy()
    genome_np = genome.cpu().numpy()
    del genome, bases
    windows = reads_np.shape[0] * (L - k + 1)
    with tempfile.TemporaryDirectory(dir=tmp_root()) as tmp:
        fa, fq, mc = os.path.join(tmp, "draft.fa"), os.path.join(tmp, "reads.fq"), os.path.join(tmp, "mult.csv")
        write_draft_fasta(fa, genome_np, starts, ends)
        write_fastq(np, fq, reads_np, bc_np, L)
        write_mult_csv(np, mc, np.unique(bc_np), mult)
        tj = os.path.join(tmp, "timing.json")
        subprocess.check_call([REF] + common_cli_args(cfg, fa) + ["-u", mc, "-t", str(threads), "-b", os.path.join(tmp, "out"),
                               "--timing-json", tj, "--map-repeats", str(args.warmup + args.steps), fq],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t = json.load(open(tj))
    runs = t["t_map_runs"][args.warmup:]
    tot_s = sum(runs)
    v = windows * len(runs) / tot_s
    sample = ("%.1f Mbp draft (%d contigs) + %d read pairs of 2x%d bp per step (same generator as the GPU arm's workload%s), -t %d, "
              "uncompressed FASTQ from tmpfs, mapping phase (readChroms: kseq parse + BX + bestContig) only, index built once"
              % (cg["genome"] / 1e6, cg["contigs"], n_pairs, L,
                 "" if cg["genome"] == cfg["genome"] else "; draft bounded to what the CPU indexes in a minute", threads))
    print(json.dumps({
        "impl": "reference", "metric": "read k-mers/s (read->contig lookup)", "value": v, "unit": "k-mers/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * tot_s / len(runs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64 (2-bit packed k-mers)",
        "data": "synthetic",
        "config": {"workload": cfg["title"] + " (bounded CPU sample: " + sample + ")", "config": cfg["name"]},
        "cpu_baseline": {"value": v, "unit": "k-mers/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------ GPU arm

def main():
    args, cfg = parse_args()
    if args.impl == "reference":
        return main_reference(args, cfg)
    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    import arcs_b200
    from arcs_b200 import merge

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    L_ = arcs_b200.load_library()
    if world > 1:
        L_.arks_bind_thread(local)  # this rank's threads and pinned buffers on the GPU's NUMA node
    numa = C.c_int(-1)
    L_.arks_device_numa_node(local, C.byref(numa))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, J, RL = cfg["k"], cfg["j"], cfg["read_len"]
    BPK = bytes_per_kmer(RL, K)

    # ---- workload: same draft on every rank (replicated table), rank-specific read shard
    genome, starts, ends = make_draft(torch, dev, cfg["genome"], cfg["contigs"], seed=1)
    draft_digest = digest(torch, genome[:min(genome.numel(), 200_000_000)])
    genome2, _, _ = make_draft(torch, dev, cfg["genome"], cfg["contigs"], seed=1)
    if not torch.equal(genome, genome2):
        raise SystemExit("bench.py: the draft generator is not deterministic")
    del genome2
    iv = contig_ends(starts, ends)
    end_bases = torch.cat([genome[s:e] for s, e, _ in iv])
    h_end_off = np.zeros(len(iv) + 1, dtype=np.uint64)
    h_end_off[1:] = np.cumsum([e - s for s, e, _ in iv])
    d_end_off = torch.from_numpy(h_end_off.astype(np.int64)).to(dev)
    d_conreci = torch.tensor([c for _, _, c in iv], dtype=torch.int32, device=dev)
    n_contigs = len(iv) // 2
    rankv = np.argsort(np.argsort(np.array([str(i + 1) for i in range(n_contigs)]))).astype(np.uint32)

    idx = arcs_b200.ArksIndex(K, int(h_end_off[-1]), device=local)
    # the kernels must run on the stream the CUDA events are recorded on: a real (non-default) torch stream,
    # made current so that the workload generation, the events and the library all use it (a null stream
    # handle would make the library fall back to its own stream, which torch's events do not see)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.current_stream().synchronize()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    idx.set_stream(stream.cuda_stream)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    idx.add_ends_device(end_bases.data_ptr(), d_end_off.data_ptr(), d_conreci.data_ptr(), h_end_off)
    ist = idx.finalize().as_dict()
    ev1.record()
    torch.cuda.synchronize()
    index_ms = ev0.elapsed_time(ev1)
    del end_bases
    if world > 1:
        merge.init_comm(idx, dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def links(mult_np):
        idx.pair_links_run(mult_np, cfg["min_mult"], cfg["max_mult"], cfg["min_reads"], ERROR_PERCENT, rankv)

    # ---- invariance workload: a fixed set of pairs, split over the ranks by barcode; the digest of the merged
    # pair-link map must be the same at every N
    inv = None
    if args.invariance_pairs > 0:
        ib, ibc, imult = make_reads(torch, dev, genome, cfg, args.invariance_pairs, seed=1234)
        keep = torch.nonzero((ibc % world) == rank).flatten()
        ib = ib.view(-1, 2 * RL)[keep].reshape(-1).contiguous()
        ibc = ibc[keep].contiguous()
        ioff = (torch.arange(2 * keep.numel() + 1, device=dev, dtype=torch.int64) * RL).to(torch.int32)
        idx.imap_clear()
        if keep.numel():
            idx.map_pairs_device(ib.data_ptr(), ioff.data_ptr(), ibc.data_ptr(), int(keep.numel()), int(ib.numel()), J, None)
        links(imult)
        if world > 1:
            merge.merge_pmap(idx)
        d0, d1 = idx.pmap_digest()
        inv = {"pairs": args.invariance_pairs, "pmap_rows": idx.pmap_size(), "pmap_digest": "%016x%016x" % (d0, d1),
               "note": "the same %d pairs at every N, dealt to the ranks by barcode; equal digests = the merged pair-link map "
                       "does not depend on the number of GPUs" % args.invariance_pairs}
        del ib, ibc, ioff, keep
        idx.imap_clear()
        idx.map_stats_reset()

    # ---- this rank's reads
    if args.strong:
        tb, tbc, mult = make_reads(torch, dev, genome, cfg, cfg["pairs"], seed=2)
        keep = torch.nonzero((tbc % world) == rank).flatten()
        bases = tb.view(-1, 2 * RL)[keep].reshape(-1).contiguous()
        barcode = tbc[keep].contiguous()
        del tb, tbc, keep
    else:
        bases, barcode, mult = make_reads(torch, dev, genome, cfg, cfg["pairs"], seed=2 + rank)
        if world > 1:  # barcode-disjoint shards: rank r's barcode ids follow rank r-1's
            nb = torch.tensor([len(mult)], device=dev, dtype=torch.int64)
            allnb = [torch.zeros_like(nb) for _ in range(world)]
            dist.all_gather(allnb, nb)
            sizes = [int(x.item()) for x in allnb]
            base_id = sum(sizes[:rank])
            barcode += base_id
            full = np.zeros(sum(sizes), dtype=np.int32)
            full[base_id:base_id + len(mult)] = mult
            mult = full
    n_pairs = int(barcode.numel())
    reads_digest = digest(torch, bases[:min(bases.numel(), 200_000_000)])
    bp = min(cfg["batch_pairs"], max(n_pairs, 1))
    assert bp * 2 * RL < 2 ** 32
    n_batches = (n_pairs + bp - 1) // bp
    # per-batch offsets are relative to the batch's first base
    off = (torch.arange(2 * bp + 1, device=dev, dtype=torch.int64) * RL).to(torch.int32)
    torch.cuda.synchronize()

    def step_device():
        for b in range(n_batches):
            a = b * bp
            n = min(bp, n_pairs - a)
            idx.map_pairs_device(bases.data_ptr() + a * 2 * RL, off.data_ptr(), barcode.data_ptr() + 4 * a, n, n * 2 * RL, J, None)

    # ---- parity sample + CPU baseline + command-line wall clock (rank 0, N = 1; not timed)
    parity, cpu, cliw = None, None, None
    if world == 1 and not args.no_cpu:
        genome_np = genome.cpu().numpy()
        threads = os.cpu_count() or 1

        def sample(n_want):
            """every pair of the first barcodes that together hold about n_want pairs, in file order"""
            cum = np.cumsum(mult.astype(np.int64) // 2)
            nb = int(np.searchsorted(cum, min(n_want, n_pairs))) + 1
            sel = torch.nonzero(barcode < nb).flatten()
            return (bases.view(-1, 2 * RL)[sel].reshape(-1, RL).cpu().numpy(), barcode[sel].cpu().numpy())

        if cfg["genome"] <= max(cfg["cpu_genome"], 200_000_000):
            s_reads, s_bc = sample(args.parity_pairs)
            parity, cpu = parity_sample(np, cfg, genome_np, starts, ends, s_reads, s_bc, mult, threads)
        else:
            # the CPU indexes ~1.6 M draft k-mers per second: a Gbp draft would take minutes.  Same generators, same
            # parameters, draft bounded to cfg.cpu_genome (as in the reference arm)
            pg = dict(cfg, genome=cfg["cpu_genome"], contigs=max(2, cfg["contigs"] * cfg["cpu_genome"] // cfg["genome"]))
            p_genome, p_starts, p_ends = make_draft(torch, dev, pg["genome"], pg["contigs"], seed=1)
            p_bases, p_bc, p_mult = make_reads(torch, dev, p_genome, pg, args.parity_pairs, seed=2)
            parity, cpu = parity_sample(np, pg, p_genome.cpu().numpy(), p_starts, p_ends, p_bases.view(-1, RL).cpu().numpy(),
                                        p_bc.cpu().numpy(), p_mult, threads)
            parity["note"] += "; draft bounded to %d Mbp for the CPU side (same generators and parameters)" % (pg["genome"] // 1_000_000)
            del p_genome, p_bases, p_bc
        if parity.get("status") not in ("ok", "oracle/_ref/arcs_ref not built"):
            print(json.dumps({"parity_sample": parity}), file=sys.stderr)
            raise SystemExit("bench.py: parity sample failed: the CUDA path and the reference's code disagree")
        if not args.no_cli:
            s_reads, s_bc = sample(args.cli_pairs)
            cliw = cli_wall(np, cfg, genome_np, starts, ends, s_reads, s_bc, mult)
        del genome_np

    # one counted pass to learn the work per step
    idx.map_stats_reset()
    step_device()
    st = idx.map_stats().as_dict()
    kmers_per_step = st["kmers_valid"] + st["kmers_invalid"]

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

    # roofline of the lookup kernels: one batch = map_groups_kernel (lane-per-read) + map_slow_kernel (the deferred
    # mates); algorithmic bytes of the batch / measured duration of the pair of launches (this rank)
    launch_ms = ms / (args.steps * n_batches)
    achieved = BPK * (kmers_per_step / n_batches) / (launch_ms / 1000.0) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    traffic = None
    tp = os.path.join(ROOT, "profiles", "map_kernel_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj.get(cfg["name"], {}).get("dram_bytes_per_launch") if cfg["name"] in tj else (
            tj.get("dram_bytes_per_launch") if cfg["name"] == "c2" else None)

    # ---- end to end: pinned host buffers through arks_map_pairs (H2D inside the timed region)
    e2e = None
    if not args.no_e2e:
        nbytes = n_pairs * 2 * RL
        hp, hb, ho = C.c_void_p(), C.c_void_p(), C.c_void_p()
        assert L_.arks_host_alloc(C.byref(hp), max(nbytes, 1)) == 0
        assert L_.arks_host_alloc(C.byref(hb), max(4 * n_pairs, 4)) == 0
        assert L_.arks_host_alloc(C.byref(ho), 4 * (2 * bp + 1)) == 0
        torch.cuda.synchronize()
        # fill the pinned buffers from the device copy (outside the timed region)
        h_bases = torch.frombuffer((C.c_uint8 * max(nbytes, 1)).from_address(hp.value), dtype=torch.uint8)
        h_bases[:nbytes].copy_(bases)
        h_bc = torch.frombuffer((C.c_uint8 * max(4 * n_pairs, 4)).from_address(hb.value), dtype=torch.int32)
        h_bc[:n_pairs].copy_(barcode)
        h_off = torch.frombuffer((C.c_uint8 * (4 * (2 * bp + 1))).from_address(ho.value), dtype=torch.int32)
        h_off.copy_(off)
        torch.cuda.synchronize()

        def step_host():
            for b in range(n_batches):
                a = b * bp
                n = min(bp, n_pairs - a)
                idx.map_pairs_raw(hp.value + a * 2 * RL, ho.value, hb.value + 4 * a, n, J)
            return idx.map_stats()  # device->host read of the step's result (counters); synchronises

        for _ in range(2):
            step_host()
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            step_host()
        e1.record()
        barrier()
        e2e_ms = max(e0.elapsed_time(e1), 0.0)
        wall_ms = (time.perf_counter() - t0) * 1000
        mine = max(e2e_ms, wall_ms)
        t = torch.tensor([mine], device=dev, dtype=torch.float64)
        per_rank = [torch.zeros_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(per_rank, t)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        else:
            per_rank = [t]
        e2e_step_s = float(t.item()) / args.steps / 1000.0
        h2d = int(nbytes + 4 * n_pairs + n_batches * 4 * (2 * bp + 1))
        e2e = {"value": float(k_all.item()) / e2e_step_s, "unit": "k-mers/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": C.sizeof(arcs_b200.MapStats) + 8, "ms_per_step": e2e_step_s * 1000,
               "h2d_gb_per_s_per_rank": [round(h2d * args.steps / (float(x.item()) / 1000.0) / 1e9, 2) for x in per_rank],
               "numa_node_of_gpu": int(numa.value), "bound_to_numa_node": world > 1 and numa.value >= 0,
               "note": "reads already parsed into pinned batch buffers (ASCII bases + offsets + barcode ids); FASTQ parsing is "
                       "measured by cli_wall / parity_sample"}
        L_.arks_host_free(hp)
        L_.arks_host_free(hb)
        L_.arks_host_free(ho)

    # ---- the job after the lookups: ONE pass over the reads, pair links on the device (ordered rows), the merge of
    # the sparse pair-link maps over NCCL (N > 1), and the export of the result to pinned host memory on rank 0
    job = None
    if not args.no_job:
        idx.imap_clear()
        step_device()
        links(mult)  # untimed: the workspaces of the pair-link stage are allocated on first use
        if world > 1:
            merge.merge_pmap(idx)
        idx.imap_clear()
        barrier()
        j0, j1, j2, j3 = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        tw0 = time.perf_counter()
        j0.record()
        step_device()
        j1.record()
        links(mult)
        j2.record()
        own_rows = idx.pmap_size()
        if world > 1:
            merge.merge_pmap(idx)
        j3.record()
        torch.cuda.synchronize()
        tw1 = time.perf_counter()
        n_rows = idx.pmap_size()
        export_ms = None
        if rank == 0 and n_rows:
            buf = C.c_void_p()
            assert L_.arks_host_alloc(C.byref(buf), n_rows * 24) == 0
            tx0 = time.perf_counter()
            got = idx.pmap_export_raw(buf.value, buf.value + 4 * n_rows, buf.value + 8 * n_rows, n_rows)
            export_ms = (time.perf_counter() - tx0) * 1000
            assert got == n_rows
            L_.arks_host_free(buf)
        d0, d1 = idx.pmap_digest()
        times = torch.tensor([j0.elapsed_time(j1), j1.elapsed_time(j2), j2.elapsed_time(j3), (tw1 - tw0) * 1000], device=dev,
                             dtype=torch.float64)
        if world > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
        pass_ms, links_ms, merge_ms, total_ms = [float(x) for x in times.tolist()]
        job = {"pass_ms": pass_ms, "pair_links_ms": links_ms, "pmap_merge_ms": merge_ms if world > 1 else None,
               "pass_plus_links_plus_merge_ms": total_ms, "pmap_export_ms": export_ms, "pmap_rows_this_rank": own_rows,
               "pmap_rows": n_rows, "pmap_digest": "%016x%016x" % (d0, d1),
               "kmers_per_s": float(k_all.item()) / (total_ms / 1000.0),
               "note": "max over ranks; pair links = pairContigs on the device incl. the radix sort into std::map order; merge = "
                       "arks_merge_pmap (NCCL all-gather of keys + one all-reduce of the dense counters); export = ordered rows to "
                       "pinned host memory on rank 0"}

    if rank == 0:
        out = {
            "metric": "read k-mers/s (read->contig lookup)", "value": value, "unit": "k-mers/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "u8/u64 (2-bit packed k-mers)",
            "data": "synthetic",
            "config": {
                "workload": "%s: synthetic %d Mbp draft (%d contigs) + %d pairs of 2x%d bp %s, k=%d, j=%.2f, -c %d -m %d-%d" % (
                    cfg["title"], cfg["genome"] // 1_000_000, cfg["contigs"], n_pairs, RL,
                    "in total, dealt to the GPUs by barcode" if args.strong else "per GPU", K, J, cfg["min_reads"], cfg["min_mult"],
                    cfg["max_mult"]),
                "config": cfg["name"],
                "l2": "inputs larger than L2 (reads %.1f GB + table %.1f GB per pass)" % (
                    n_pairs * 2 * RL / 1e9, ist["recorded"] * 2 * 32 / 1e9),
                "batches_per_step": n_batches, "kmers_per_step_per_gpu": kmers_per_step, "table_keys": ist["recorded"],
                "draft_digest": "%016x" % draft_digest, "reads_digest_rank0": "%016x" % reads_digest,
                "generator": "seeded torch generators (draft seed 1, reads seed 2 + rank); the draft is generated twice and "
                             "compared",
                "index_build_ms": index_ms,
                "parallelism": "barcode-sharded reads x%d, replicated k-mer table" % world,
            },
            "clocks": clk, "gpu_launches": int(gpu_launches), "e2e": e2e, "job": job, "invariance": inv,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "kernel": "map_groups_kernel (lane-per-read) + map_slow_kernel (deferred mates); one batch of %d pairs" % bp,
                         "bytes_per_kmer": BPK, "peak_source": peak_src, "launch_ms": launch_ms},
            "cpu_baseline": cpu, "parity_sample": parity, "cli_wall": cliw,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
