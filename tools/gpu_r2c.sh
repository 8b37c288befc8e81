#!/usr/bin/env bash
# tests, sanitizers, job timing, slow-kernel register A/B on c5
mkdir -p gpurun_out
O=gpurun_out/r2c
(timeout 1500 python -m pytest tests -m gpu -q) > $O.pytest.log 2>&1
echo "pytest rc=$?"; tail -12 $O.pytest.log
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "demo_golden_gpu or chimeric or more_than_32 or pair_links" > $O.$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" $O.$tool.log | tail -4
done
ARKS_TIMING=1 timeout 900 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu --no-e2e > $O.bench_c2.json 2> $O.bench_c2.err
grep -E "arks_pair_links" $O.bench_c2.err | tail -14
python -c "
import json;d=json.loads(open('$O.bench_c2.json').read().strip().splitlines()[-1]);print(d['value'],d['job'])"
export AB_CONFIG=c5 AB_PAIRS=3000000
SKIP="tests memcheck full ncu" BUILDS="-DARKS_MAP_MIN_BLOCKS=4;-DARKS_MAP_MIN_BLOCKS=3;-DARKS_MAP_MIN_BLOCKS=5 -DARKS_SEEDS=4" VARIANTS="X=1" bash tools/gpu_round.sh r2c_ab
