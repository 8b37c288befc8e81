#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out/r2d
(timeout 1500 python -m pytest tests -m gpu -q -x) > $O.pytest.log 2>&1
echo "pytest rc=$?"; tail -6 $O.pytest.log
for v in 1 0; do
  ARKS_MFILTER=$v timeout 600 python bench.py --config c2 --pairs 6250000 --steps 3 --warmup 3 --no-cpu --no-e2e --no-job --invariance-pairs 0 > $O.c2_mf$v.json 2> $O.c2_mf$v.err
  python -c "
import json;d=json.loads(open('$O.c2_mf$v.json').read().strip().splitlines()[-1]);print('c2 mfilter=$v value=%.4e launch_ms=%.3f frac=%.3f'%(d['value'],d['roofline']['launch_ms'],d['roofline']['frac']))" || tail -3 $O.c2_mf$v.err
done
for v in 1 0; do
  ARKS_MFILTER=$v timeout 900 python bench.py --config c2 --genome 2000000000 --contigs 200000 --pairs 3125000 --steps 3 --warmup 3 --no-cpu --no-e2e --no-job --invariance-pairs 0 > $O.2g_mf$v.json 2> $O.2g_mf$v.err
  python -c "
import json;d=json.loads(open('$O.2g_mf$v.json').read().strip().splitlines()[-1]);print('2Gbp mfilter=$v value=%.4e launch_ms=%.3f frac=%.3f index_ms=%.1f keys=%d'%(d['value'],d['roofline']['launch_ms'],d['roofline']['frac'],d['config']['index_build_ms'],d['config']['table_keys']))" || tail -3 $O.2g_mf$v.err
done
timeout 900 python bench.py --config c2 > $O.bench_c2.json 2> $O.bench_c2.err; tail -c 600 $O.bench_c2.err
python -c "
import json;d=json.loads(open('$O.bench_c2.json').read().strip().splitlines()[-1]);print('c2 full value=%.4e e2e=%.4e frac=%.3f parity=%s'%(d['value'],d['e2e']['value'],d['roofline']['frac'],d['parity_sample']['status'])); print(d['cli_wall'])"
timeout 900 python bench.py --config c3 > $O.bench_c3.json 2> $O.bench_c3.err; tail -c 600 $O.bench_c3.err
python -c "
import json;d=json.loads(open('$O.bench_c3.json').read().strip().splitlines()[-1]);print('c3 full value=%.4e e2e=%.4e frac=%.3f parity=%s'%(d['value'],d['e2e']['value'],d['roofline']['frac'],d['parity_sample']['status']))"
