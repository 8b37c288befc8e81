import os, subprocess, sys, tempfile, time
sys.path.insert(0, '.')
import numpy as np
import bench
tmp = tempfile.mkdtemp(prefix='abgz_')
fa, fq, mult, windows = bench.write_cpu_sample(np, tmp, 10_000_000, 2_000_000, 7)
subprocess.check_call(['gzip', '-1', '-k', '-f', fq])
common = ['-f', fa, '-k', '60', '-j', '0.55', '-c', '5', '-m', '50-10000', '-e', '30000', '-z', '500', '-r', '0.05']
for rep in range(2):
    for mode, env in (('fast_inflate', {}), ('zlib', {'ARKS_ZLIB': '1'})):
        e = dict(os.environ); e.update(env)
        t = time.time()
        out = subprocess.run(['arcs_b200/bin/arcs', '--arks', '-v'] + common + ['-b', os.path.join(tmp, mode), fq + '.gz'], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
        dt = time.time() - t
        line = [l for l in out.splitlines() if l.startswith('GPU mapping')]
        print(mode, '%.2f s' % dt, line[0][:60] if line else out[-300:])
    t = time.time(); subprocess.run(['arcs_b200/bin/inflate_check', 'fast', fq + '.gz', '1048576', 'q']); 
    subprocess.run(['arcs_b200/bin/inflate_check', 'zlib', fq + '.gz', '1048576', 'q'])
