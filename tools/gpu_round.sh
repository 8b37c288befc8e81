#!/usr/bin/env bash
# tools/gpu_round.sh <tag> -- one GPU-box visit: parity tests, memcheck of the smoke pass, A/B bench
# runs over run-time knobs (VARIANTS, space separated VAR=val) and compile-time knobs (BUILDS,
# ';'-separated nvcc flag sets), the full bench line, the ncu launch list and one full ncu capture.
# Sections are skipped with SKIP="tests memcheck full ncu".  Everything lands in gpurun_out/<tag>.*
tag=${1:-r}
mkdir -p gpurun_out
O=gpurun_out/$tag
skip() { case " $SKIP " in *" $1 "*) return 0;; esac; return 1; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O.gpu.txt 2>&1
prc=0
if ! skip tests; then
  (timeout 1200 python -m pytest tests -m gpu -x -q) > $O.pytest.log 2>&1
  prc=$?
  echo "pytest rc=$prc"; tail -5 $O.pytest.log
fi
if ! skip memcheck; then
  timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O.memcheck.log 2>&1
  echo "memcheck rc=$?"; tail -4 $O.memcheck.log
fi
short_bench() { # label, outfile, env...
  local label=$1 out=$2; shift 2
  env "$@" timeout 600 python bench.py --pairs ${AB_PAIRS:-6250000} --steps 3 --warmup 3 --no-cpu --no-e2e --no-job --invariance-pairs 0 ${AB_CONFIG:+--config $AB_CONFIG} > $out 2> $out.err
  python - "$label" $out <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print('%-44s value=%.3e k-mers/s  launch_ms=%.3f  frac=%.3f' % (sys.argv[1], d['value'], d['roofline']['launch_ms'], d['roofline']['frac']))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
i=0
for v in ${VARIANTS-X=1}; do
  short_bench "$v" $O.ab$i.json $v
  i=$((i+1))
done
if [ -n "$BUILDS" ]; then
  IFS=';' read -ra BL <<< "$BUILDS"
  for b in "${BL[@]}"; do
    make -s -C arcs_b200/csrc clean >/dev/null
    make -s -C arcs_b200/csrc EXTRA="$b" 2>&1 | grep -A2 'map_groups_kernelILi2' | grep -E 'Used|spill' | sed 's/ptxas info *: //' | tr '\n' ' '
    echo
    short_bench "BUILD[$b]" $O.build$i.json X=1
    i=$((i+1))
  done
  make -s -C arcs_b200/csrc clean >/dev/null; make -s -C arcs_b200/csrc ${FINAL_EXTRA:+EXTRA="$FINAL_EXTRA"} >/dev/null 2>&1
fi
if ! skip full && { [ "$prc" = "0" ] || [ -n "$FORCE_FULL" ]; }; then
  timeout 900 python bench.py > $O.bench.json 2> $O.bench.err
  tail -c 3000 $O.bench.json
fi
if ! skip ncu; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O.launches.csv \
    python bench.py --pairs 3125000 --steps 2 --warmup 3 --no-cpu --no-e2e --no-job --invariance-pairs 0 ${AB_CONFIG:+--config $AB_CONFIG} > $O.launches.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:map_ -s 4 -c 2 -o $O.map_full -f \
    python bench.py --pairs 3125000 --steps 1 --warmup 1 --no-cpu --no-e2e --no-job --invariance-pairs 0 ${AB_CONFIG:+--config $AB_CONFIG} > $O.ncu_full.log 2>&1
fi
ls gpurun_out | tr '\n' ' '
