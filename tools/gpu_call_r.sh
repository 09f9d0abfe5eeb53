#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for wl in C2 C4; do
  timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-pipeline --profile > gpurun_out/r02t.json 2> gpurun_out/r02t_breakdown.txt
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02t.json").read().strip().splitlines()[-1])
print("$wl ms/traj", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["per_bin_violation"], d["parity"]["rms_rel_err"], "frac", round(d["roofline"]["frac"],4))
PY
  grep -E "dw_patch|la_weff|la_combine|tail_merge|attn_fwd" gpurun_out/r02t_breakdown.txt
done
