#!/usr/bin/env bash
# tools/sweep.sh -- rebuild the library with different compile-time knobs on the GPU box and
# time the lookup kernel on a reduced configs[1] workload.  usage: tools/sweep.sh "<EXTRA flags>" ...
set -e
for extra in "$@"; do
  make -s -C arcs_b200/csrc clean >/dev/null
  make -s -C arcs_b200/csrc EXTRA="$extra" 2>&1 | grep -A2 'map_pairs_kernelILi2' | grep -E 'Used' | sed 's/ptxas info *: //' || true
  python bench.py --pairs 6250000 --steps 3 --warmup 2 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('EXTRA=[$extra] value=%.3e kmers/s ms_per_step=%.2f frac=%.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
done
make -s -C arcs_b200/csrc clean >/dev/null; make -s -C arcs_b200/csrc >/dev/null 2>&1
