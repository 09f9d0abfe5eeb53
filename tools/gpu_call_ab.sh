#!/bin/bash
# same-box A/B of several builds (ab/<commit>/ holds an archived tree with its own library), two rounds, interleaved
mkdir -p gpurun_out
for round in 1 2; do
for tag in 5a440b2 9379bf6 fa2e987 HEAD; do
  if [ $tag = HEAD ]; then dir=.; else dir=ab/$tag; fi
  (cd $dir && timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipeline 2>/dev/null) | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag round $round: ms/traj', round(d['ms_per_step'],2), 'clk', d['clocks']['sm_mhz'], 'e2e', round(d['e2e']['value']))"
done; done
