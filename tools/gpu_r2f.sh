#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out/r2f
(timeout 900 python -m pytest tests/test_cli_gpu.py tests/test_zz_cut_gpu.py tests/test_gpu_merge.py -m gpu -q -x) > $O.pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O.pytest.log
ARKS_TIMING=1 timeout 1200 python bench.py --config c2 --genome 3000000000 --contigs 300000 --pairs 3125000 --steps 3 --warmup 3 --no-cpu --no-e2e --no-job --invariance-pairs 0 > $O.3g.json 2> $O.3g.err
python -c "
import json;d=json.loads(open('$O.3g.json').read().strip().splitlines()[-1]);print('3Gbp value=%.4e launch_ms=%.3f frac=%.3f index_ms=%.1f keys=%d'%(d['value'],d['roofline']['launch_ms'],d['roofline']['frac'],d['config']['index_build_ms'],d['config']['table_keys']))" || tail -5 $O.3g.err
ARKS_TIMING=1 timeout 1500 python tools/big_run.py --genome 3000000000 --contigs 300000 --pairs 20000000 --gpus 1 > $O.big1.json 2> $O.big1.err
grep -E "^\{" $O.big1.err | cut -c 1-900; cut -c 1-300 $O.big1.json
