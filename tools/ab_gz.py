#!/usr/bin/env python
"""A/B of the gzip ingest on the GPU box: the CLI on the same .fq.gz with the multi-threaded decoder
(par_inflate.h, default), the single-threaded one (ARKS_GZ_THREADS=1) and zlib (ARKS_ZLIB=1), plus the bare
decoders.  Results are appended to gpurun_out/ab_gz.txt as they come.   python tools/ab_gz.py [pairs]"""
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, '.')
import numpy as np

import bench

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
os.makedirs('gpurun_out', exist_ok=True)
log = open('gpurun_out/ab_gz.txt', 'a')


def say(*a):
    msg = ' '.join(str(x) for x in a)
    print(msg, flush=True)
    log.write(msg + '\n')
    log.flush()


tmp = tempfile.mkdtemp(prefix='abgz_')
import torch  # noqa: E402

dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
cfg = dict(bench.CONFIGS["c2"], genome=10_000_000, contigs=1000, name="c2")
genome, starts, ends = bench.make_draft(torch, dev, cfg["genome"], cfg["contigs"], seed=7)
bases, barcode, _ = bench.make_reads(torch, dev, genome, cfg, pairs, seed=8)
fa, fq = os.path.join(tmp, "draft.fa"), os.path.join(tmp, "reads.fq")
bench.write_draft_fasta(fa, genome.cpu().numpy(), starts, ends)
bench.write_fastq(np, fq, bases.view(-1, cfg["read_len"]).cpu().numpy(), barcode.cpu().numpy(), cfg["read_len"])
subprocess.check_call(['gzip', '-1', '-f', fq])
gz = fq + '.gz'
say('sample: %d pairs, %.2f GB compressed' % (pairs, os.path.getsize(gz) / 1e9))
for mode in ('par8', 'fast', 'zlib'):
    p = subprocess.run(['arcs_b200/bin/inflate_check', mode, gz, '1048576', 'q'], stderr=subprocess.PIPE, text=True)
    say('decoder', mode, p.stderr.strip().splitlines()[-1])
common = ['-f', fa, '-k', '60', '-j', '0.55', '-c', '5', '-m', '50-10000', '-e', '30000', '-z', '500', '-r', '0.05']
for mode, env in (('par_inflate(default)', {}), ('fast_inflate(1 thread)', {'ARKS_GZ_THREADS': '1'}), ('zlib', {'ARKS_ZLIB': '1'})):
    e = dict(os.environ)
    e.update(env)
    t = time.time()
    out = subprocess.run(['arcs_b200/bin/arcs', '--arks', '-v'] + common + ['-b', os.path.join(tmp, 'o'), gz], env=e, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True).stdout
    line = [ln for ln in out.splitlines() if ln.startswith('GPU mapping')]
    say('cli', mode, '%.2f s total;' % (time.time() - t), line[0][:70] if line else out[-200:])
