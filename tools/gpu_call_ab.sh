#!/bin/bash
# same-box A/B: ab/<tag>/ holds an archived tree of an older commit with its own library; interleaved runs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_gpu.py tests/test_scale_gpu.py tests/test_taps_gpu.py tests/test_libritts_gpu.py -x -q 2>&1 | tail -2
for round in 1 2; do
for tag in prev HEAD; do
  if [ $tag = HEAD ]; then dir=.; else dir=ab/$tag; fi
  (cd $dir && timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipeline --profile 2>gpurun_out_bd.txt) | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag round $round: ms/traj', round(d['ms_per_step'],2), 'clk', d['clocks']['sm_mhz'], 'frac', round(d['roofline']['frac'],4), 'parity', d['parity']['per_bin_violation'])"
  (cd $dir && grep -E "gn_apply|gn_final" gpurun_out_bd.txt | awk '{printf "    %s %s ms\n", $1, $4}')
done; done
