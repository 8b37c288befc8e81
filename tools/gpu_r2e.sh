#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out/r2e
export AB_CONFIG=c2 AB_PAIRS=6250000
SKIP="tests memcheck full ncu" BUILDS="-DARKS_CHAIN3;-DARKS_PREFETCH_NEXT;-DARKS_CHAIN3 -DARKS_PREFETCH_NEXT" VARIANTS="X=1" bash tools/gpu_round.sh r2e_ab
ARKS_TIMING=1 timeout 1200 python bench.py --config c2 --genome 3000000000 --contigs 300000 --pairs 3125000 --steps 3 --warmup 3 --no-cpu --no-e2e --no-job --invariance-pairs 0 > $O.3g.json 2> $O.3g.err
tail -c 1200 $O.3g.err
python -c "
import json;d=json.loads(open('$O.3g.json').read().strip().splitlines()[-1]);print('3Gbp value=%.4e launch_ms=%.3f frac=%.3f index_ms=%.1f keys=%d'%(d['value'],d['roofline']['launch_ms'],d['roofline']['frac'],d['config']['index_build_ms'],d['config']['table_keys']))"
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
ARKS_TIMING=1 timeout 1500 python tools/big_run.py --genome 3000000000 --contigs 300000 --pairs 20000000 --gpus 1 > $O.big1.json 2> $O.big1.err
tail -c 2500 $O.big1.err; cat $O.big1.json | cut -c 1-1500
