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
fa, fq, mult, windows = bench.write_cpu_sample(np, tmp, 10_000_000, pairs, 7)
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
