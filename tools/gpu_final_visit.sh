#!/usr/bin/env bash
# final one-GPU visit of the round: the whole GPU suite, sanitizers, smoke, default bench lines, gzip input at scale
mkdir -p gpurun_out
O=gpurun_out/final
(timeout 1500 python -m pytest tests -m gpu -q -x) > $O.pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O.pytest.log
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "demo_golden_gpu or chimeric or more_than_32 or pair_links or garbage_keys" > $O.$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" $O.$tool.log | tail -3
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O.smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O.smoke.log
timeout 900 python bench.py > $O.bench_c2.json 2> $O.bench_c2.err; echo "bench rc=$?"; tail -c 300 $O.bench_c2.err
python -c "
import json;d=json.loads(open('$O.bench_c2.json').read().strip().splitlines()[-1]);print('c2 value=%.4e e2e=%.4e frac=%.3f parity=%s job=%.2f ms cli=%.2f s'%(d['value'],d['e2e']['value'],d['roofline']['frac'],d['parity_sample']['status'],d['job']['pass_plus_links_plus_merge_ms'],d['cli_wall']['wall_s']))"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O.bench_ref.json 2> $O.bench_ref.err; echo "ref rc=$?"; cut -c 1-400 $O.bench_ref.json
timeout 900 python tools/big_run.py --genome 500000000 --contigs 50000 --pairs 10000000 --gpus 1 --gzip > $O.big_gz.json 2> $O.big_gz.err
grep -E "^\{" $O.big_gz.err | cut -c 1-900
