#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipeline --profile > gpurun_out/r02t.json 2> gpurun_out/r02t_breakdown.txt
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02t.json").read().strip().splitlines()[-1])
print("ms/traj", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], "parity", d["parity"], "frac", round(d["roofline"]["frac"],4))
PY
done
DEXB_PDL=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipeline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no PDL: ms/traj', round(d['ms_per_step'],2), 'clk', d['clocks']['sm_mhz'])"
