#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_gpu.py tests/test_taps_gpu.py -x -q 2>&1 | tail -2
for pp in 8 0 4 16 8 0; do
  DEXB_GN_PIPE=$pp timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile > gpurun_out/r02l_p$pp.json 2> gpurun_out/r02l_p${pp}_breakdown.txt
  python - <<PY
import json
d=json.load(open("gpurun_out/r02l_p$pp.json"))
print("pipe=$pp: ms/traj", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["per_bin_violation"])
PY
  grep "gn_apply" gpurun_out/r02l_p${pp}_breakdown.txt
done
