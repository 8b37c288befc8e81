#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out/r2l
idx() { # label
  for i in 1 2 3; do
    timeout 600 python bench.py --config c2 --pairs 1562500 --steps 1 --warmup 3 --no-cpu --no-e2e --no-job --invariance-pairs 0 > $O.tmp.json 2> $O.tmp.err
    python -c "
import json;d=json.loads(open('$O.tmp.json').read().strip().splitlines()[-1]);print('$1 run $i: index_ms=%.2f keys=%d value=%.3e'%(d['config']['index_build_ms'],d['config']['table_keys'],d['value']))" || tail -3 $O.tmp.err
  done
}
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -q -x) > $O.pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O.pytest.log
idx "minblocks3"
make -s -C arcs_b200/csrc clean >/dev/null; make -s -C arcs_b200/csrc EXTRA="-DARKS_INSERT_MIN_BLOCKS=4" > /dev/null 2>&1
idx "minblocks4"
make -s -C arcs_b200/csrc clean >/dev/null; make -s -C arcs_b200/csrc EXTRA="-DARKS_INSERT_MIN_BLOCKS=2" > /dev/null 2>&1
idx "minblocks2"
make -s -C arcs_b200/csrc clean >/dev/null; make -s -C arcs_b200/csrc > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"insert_kernel|finalize_kernel|uniq_mask|inv_mask|walk_kernel" -c 6 --csv --log-file $O.index_launches.csv python bench.py --config c2 --pairs 1562500 --steps 1 --warmup 1 --no-cpu --no-e2e --no-job --invariance-pairs 0 > /dev/null 2>&1
grep -E "insert|finalize|uniq|inv_mask|walk" $O.index_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c 1-120
timeout 1200 python bench.py --config c2 --genome 3000000000 --contigs 300000 --pairs 3125000 --steps 2 --warmup 3 --no-cpu --no-e2e --no-job --invariance-pairs 0 > $O.3g.json 2> $O.3g.err
python -c "
import json;d=json.loads(open('$O.3g.json').read().strip().splitlines()[-1]);print('3Gbp value=%.4e index_ms=%.1f'%(d['value'],d['config']['index_build_ms']))" || tail -5 $O.3g.err
