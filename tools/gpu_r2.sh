#!/usr/bin/env bash
# tools/gpu_r2.sh <tag> -- round-2 GPU visit: parity tests, then the bench lines of the three configs
tag=${1:-r2}
mkdir -p gpurun_out
O=gpurun_out/$tag
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O.gpu.txt 2>&1
nproc >> $O.gpu.txt; free -g | head -2 >> $O.gpu.txt; lscpu | grep -E "Model name|Socket|NUMA" >> $O.gpu.txt
if [ -z "$SKIP_TESTS" ]; then
  (timeout ${TEST_TIMEOUT:-1500} python -m pytest tests -m gpu -q ${PYTEST_ARGS}) > $O.pytest.log 2>&1
  echo "pytest rc=$?"; tail -15 $O.pytest.log
fi
for c in ${CONFIGS-c2 c3 c5}; do
  timeout 1200 python bench.py --config $c ${BENCH_ARGS} > $O.bench_$c.json 2> $O.bench_$c.err
  echo "bench $c rc=$?"; tail -c 1500 $O.bench_$c.err; python - $O.bench_$c.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('value=%.3e e2e=%.3e frac=%.3f launch_ms=%.3f' % (d['value'], (d.get('e2e') or {}).get('value') or 0, d['roofline']['frac'], d['roofline']['launch_ms']))
    print('job', d.get('job')); print('parity', (d.get('parity_sample') or {}).get('status'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    print('cli', d.get('cli_wall')); print('inv', d.get('invariance')); print('cfg', d['config'])
except Exception as e:
    print('FAILED', e)
PY
done
