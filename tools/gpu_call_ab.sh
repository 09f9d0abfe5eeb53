#!/bin/bash
# Same-box A/B of two builds (boxes differ by up to 4 % and the power cap moves the SM clock with the load, so only interleaved runs on
# one box are used for attribution).  Prepare the older build next to the tree, then run this under gpurun:
#   mkdir -p ab/prev && git archive <commit> dex-tts_b200 bench.py oracle tests/parity.py tests/conftest.py BASELINE.json \
#       profiles/r02_ncu_traffic.json include | tar -x -C ab/prev && cp MEASURED_PEAKS.json ab/prev/ && bash ab/prev/dex-tts_b200/build.sh
#   gpurun --timeout 900 -- 'bash tools/gpu_call_ab.sh prev'
# (ab/ is git-ignored; delete it afterwards -- it travels to the box with every gpurun call.)
TAGS="${*:-prev} HEAD"
for round in 1 2; do
for tag in $TAGS; do
  if [ $tag = HEAD ]; then dir=.; else dir=ab/$tag; fi
  (cd $dir && timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipeline --profile 2>/tmp/ab_breakdown.txt) | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag round $round: ms/traj', round(d['ms_per_step'],2), 'clk', d['clocks']['sm_mhz'], 'frac', round(d['roofline']['frac'],4), 'parity', d['parity']['per_bin_violation'])"
  head -8 /tmp/ab_breakdown.txt | awk '{printf "    %s %s ms\n", $1, $4}'
done; done
