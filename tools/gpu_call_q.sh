#!/bin/bash
mkdir -p gpurun_out
for w in 2 1 3 4; do
  DEXB_LA_WAVES=$w timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipeline --profile > gpurun_out/r02q_w$w.json 2> gpurun_out/r02q_w${w}_breakdown.txt
  python - <<PY
import json
d=json.load(open("gpurun_out/r02q_w$w.json"))
print("la_waves=$w: ms/traj", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["per_bin_violation"])
PY
  grep -E "attn_fwd|la_combine|la_weff|la.apply" gpurun_out/r02q_w${w}_breakdown.txt
done
