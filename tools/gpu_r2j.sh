#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out/r2j
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_cli_gpu.py tests/test_gpu_merge.py -m gpu -q -x) > $O.pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O.pytest.log
for i in 1 2; do
timeout 600 python bench.py --config c2 --pairs 3125000 --steps 2 --warmup 3 --no-cpu --no-e2e --no-job --invariance-pairs 0 > $O.c2.json 2> $O.c2.err
python -c "
import json;d=json.loads(open('$O.c2.json').read().strip().splitlines()[-1]);print('c2 value=%.4e index_ms=%.2f keys=%d'%(d['value'],d['config']['index_build_ms'],d['config']['table_keys']))" || tail -3 $O.c2.err
done
timeout 1200 python bench.py --config c2 --genome 3000000000 --contigs 300000 --pairs 3125000 --steps 3 --warmup 3 --no-cpu --no-e2e --no-job --invariance-pairs 0 > $O.3g.json 2> $O.3g.err
python -c "
import json;d=json.loads(open('$O.3g.json').read().strip().splitlines()[-1]);print('3Gbp value=%.4e launch_ms=%.3f frac=%.3f index_ms=%.1f keys=%d'%(d['value'],d['roofline']['launch_ms'],d['roofline']['frac'],d['config']['index_build_ms'],d['config']['table_keys']))" || tail -5 $O.3g.err
timeout 1200 python bench.py --config c2 --genome 2000000000 --contigs 200000 --pairs 3125000 --steps 3 --warmup 3 --no-cpu --no-e2e --no-job --invariance-pairs 0 > $O.2g.json 2> $O.2g.err
python -c "
import json;d=json.loads(open('$O.2g.json').read().strip().splitlines()[-1]);print('2Gbp value=%.4e launch_ms=%.3f frac=%.3f index_ms=%.1f keys=%d'%(d['value'],d['roofline']['launch_ms'],d['roofline']['frac'],d['config']['index_build_ms'],d['config']['table_keys']))" || tail -5 $O.2g.err
