#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_gpu.py -x -q 2>&1 | tail -3
for r in 1 0 1 0; do
  DEXB_GN_REVERSE=$r timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile > gpurun_out/r02f_r$r.json 2> gpurun_out/r02f_r${r}_breakdown.txt
  python - <<PY
import json
d=json.load(open("gpurun_out/r02f_r$r.json"))
print("reverse=$r: ms/traj", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["per_bin_violation"])
PY
  grep "gn_apply\|b1.conv\|b2.conv" gpurun_out/r02f_r${r}_breakdown.txt
done
