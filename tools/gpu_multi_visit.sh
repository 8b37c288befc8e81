#!/usr/bin/env bash
# N-GPU visit (gpurun --gpus N): NCCL merge tests, CLI invariance, the bench line at N, the command line at 1 and N GPUs
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out/r2g_$N
nvidia-smi --query-gpu=index,name --format=csv > $O.gpu.txt; nproc >> $O.gpu.txt; lscpu | grep -E "Model name|Socket|NUMA" >> $O.gpu.txt; nvidia-smi topo -m >> $O.gpu.txt 2>&1
(timeout 900 python -m pytest tests/test_gpu_merge.py tests/test_cli_gpu.py -m gpu -q -k "nccl or invariant or gpu-ids") > $O.pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O.pytest.log
for n in 1 $N; do
  if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511"; fi
  timeout 900 $L bench.py --gpus $n --steps 5 --warmup 3 ${BENCH_EXTRA---no-cpu} > $O.bench_n$n.json 2> $O.bench_n$n.err
  python - $O.bench_n$n.json <<'PY' || tail -5 $O.bench_n$n.err
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('N=%d value=%.4e e2e=%.4e (%.1f ms) frac=%.3f' % (d['n_gpus'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac']))
print(' h2d/rank', d['e2e'].get('h2d_gb_per_s_per_rank'), 'numa', d['e2e'].get('numa_node_of_gpu'))
print(' job', {k:v for k,v in d['job'].items() if k!='note'})
print(' inv', d['invariance']['pmap_digest'], d['invariance']['pmap_rows'])
PY
done
timeout 1500 python tools/big_run.py --genome ${BIG_GENOME:-500000000} --contigs ${BIG_CONTIGS:-50000} --pairs ${BIG_PAIRS:-10000000} --gpus 1,$N > $O.big.json 2> $O.big.err
grep -E "^\{" $O.big.err | cut -c 1-900; python -c "
import json;d=json.load(open('$O.big.json'));print('identical across runs:',d['outputs_identical_across_runs'],'gen_s',d['generate_s'])"
