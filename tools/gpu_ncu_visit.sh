#!/usr/bin/env bash
# ncu evidence for round 2 (one GPU): launch list of the bench command (c2), full captures of the lookup kernels on c2
# and c5 (one batch each) and of the index-build kernels; the complete default bench lines of c2 / c3 / c5
mkdir -p gpurun_out
O=gpurun_out/ncu
C2="python bench.py --config c2 --pairs 3125000 --steps 2 --warmup 3 --no-cpu --no-e2e --no-job --invariance-pairs 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O.launches_c2.csv $C2 > $O.launches_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:map_ -s 4 -c 2 -o $O.map_c2 -f \
  python bench.py --config c2 --pairs 3125000 --steps 1 --warmup 1 --no-cpu --no-e2e --no-job --invariance-pairs 0 > $O.ncu_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"insert_kernel|finalize_kernel|uniq_mask|bloom_build" -c 4 -o $O.index_c2 -f \
  python bench.py --config c2 --pairs 1562500 --steps 1 --warmup 1 --no-cpu --no-e2e --no-job --invariance-pairs 0 > $O.ncu_index.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:map_ -s 4 -c 2 -o $O.map_c5 -f \
  python bench.py --config c5 --pairs 2000000 --steps 1 --warmup 1 --no-cpu --no-e2e --no-job --invariance-pairs 0 > $O.ncu_c5.log 2>&1
ls -la gpurun_out/ncu*
for c in c2 c3 c5; do
  timeout 1200 python bench.py --config $c > $O.bench_$c.json 2> $O.bench_$c.err
  echo "bench $c rc=$?"; tail -c 300 $O.bench_$c.err
done
