#!/usr/bin/env bash
# 8-GPU visit: NCCL merge + CLI invariance tests at 2/4/8, the bench line at N=8, the command line at 1 and 8 GPUs on a
# human-scale draft
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out/r2k_$N
nvidia-smi --query-gpu=index,name --format=csv > $O.gpu.txt; nproc >> $O.gpu.txt; free -g | head -2 >> $O.gpu.txt; lscpu | grep -E "Model name|Socket|NUMA" >> $O.gpu.txt; nvidia-smi topo -m >> $O.gpu.txt 2>&1
(timeout 400 python -m pytest tests/test_gpu_merge.py tests/test_cli_gpu.py -m gpu -q -k "nccl or invariant") > $O.pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O.pytest.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > $O.bench_n$N.json 2> $O.bench_n$N.err
python - $O.bench_n$N.json <<'PY' || tail -5 $O.bench_n$N.err
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('N=%d value=%.4e e2e=%.4e (%.1f ms) frac=%.3f' % (d['n_gpus'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac']))
print(' h2d/rank', d['e2e'].get('h2d_gb_per_s_per_rank'), 'numa', d['e2e'].get('numa_node_of_gpu'))
print(' job', {k:v for k,v in d['job'].items() if k!='note'})
print(' inv', d['invariance']['pmap_digest'], d['invariance']['pmap_rows'])
PY
timeout 560 python tools/big_run.py --genome ${BIG_GENOME:-3000000000} --contigs ${BIG_CONTIGS:-300000} --pairs ${BIG_PAIRS:-30000000} --gpus 1,$N > $O.big.json 2> $O.big.err
grep -E "^\{" $O.big.err | cut -c 1-1100; python -c "
import json;d=json.load(open('$O.big.json'));print('identical across runs:',d['outputs_identical_across_runs'],'gen_s',d['generate_s'])" || tail -5 $O.big.err
