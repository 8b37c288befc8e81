#!/usr/bin/env bash
# tools/gpu_round.sh <tag> -- one GPU-box visit: parity tests, memcheck of the smoke pass, A/B bench
# runs over the run-time knobs, the full bench line, the ncu launch list and one full ncu capture.
# Everything lands in gpurun_out/<tag>.*
tag=${1:-r}
mkdir -p gpurun_out
O=gpurun_out/$tag
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O.gpu.txt 2>&1
(timeout 1200 python -m pytest tests -m gpu -x -q) > $O.pytest.log 2>&1
prc=$?
echo "pytest rc=$prc"; tail -5 $O.pytest.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O.memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 $O.memcheck.log
i=0
for v in ${VARIANTS:-X=1 ARKS_LANE_GENERAL=0 ARKS_BLOOM_BITS=0 ARKS_BLOOM_BITS=6 ARKS_BLOOM_BITS=12}; do
  env $v timeout 600 python bench.py --pairs ${AB_PAIRS:-6250000} --steps 3 --warmup 3 --no-cpu --no-e2e > $O.ab$i.json 2> $O.ab$i.err
  python - "$v" $O.ab$i.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print('%-24s value=%.3e k-mers/s  launch_ms=%.3f  frac=%.3f' % (sys.argv[1], d['value'], d['roofline']['launch_ms'], d['roofline']['frac']))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
  i=$((i+1))
done
if [ "$prc" = "0" ] || [ -n "$FORCE_FULL" ]; then
  timeout 900 python bench.py > $O.bench.json 2> $O.bench.err
  tail -c 3000 $O.bench.json
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O.launches.csv \
  python bench.py --pairs 3125000 --steps 2 --warmup 3 --no-cpu --no-e2e > $O.launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:map_ -s 4 -c 2 -o $O.map_full -f \
  python bench.py --pairs 3125000 --steps 1 --warmup 1 --no-cpu --no-e2e > $O.ncu_full.log 2>&1
ls -la gpurun_out | tail -20
