#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out/r2n
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_zz_cut_gpu.py tests/test_gpu_wallclock.py -m gpu -q -x) > $O.pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O.pytest.log
for c in c2 c3 c5; do
  P=6250000; [ $c = c5 ] && P=3000000; [ $c = c3 ] && P=5000000
  timeout 900 python bench.py --config $c --pairs $P --steps 3 --warmup 3 --no-cpu --no-e2e --no-job --invariance-pairs 0 > $O.$c.json 2> $O.$c.err
  python -c "
import json;d=json.loads(open('$O.$c.json').read().strip().splitlines()[-1]);print('$c value=%.4e launch_ms=%.3f frac=%.3f'%(d['value'],d['roofline']['launch_ms'],d['roofline']['frac']))" || tail -3 $O.$c.err
done
timeout 600 python bench.py --config c2 --pairs 6250000 --batch-pairs 3125000 --steps 3 --warmup 3 --no-cpu --no-e2e --no-job --invariance-pairs 0 > $O.c2b.json 2> $O.c2b.err
python -c "
import json;d=json.loads(open('$O.c2b.json').read().strip().splitlines()[-1]);print('c2 batch 3.125M value=%.4e launch_ms=%.3f frac=%.3f'%(d['value'],d['roofline']['launch_ms'],d['roofline']['frac']))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:map_slow -s 2 -c 1 -o $O.slow_c5 -f \
  python bench.py --config c5 --pairs 2000000 --steps 1 --warmup 1 --no-cpu --no-e2e --no-job --invariance-pairs 0 > $O.ncu_c5.log 2>&1
