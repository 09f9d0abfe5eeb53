#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_gpu.py -x -q 2>&1 | tail -3
for cfg in "0 2" "1 2" "2 2" "3 2" "4 2"; do
  set -- $cfg
  DEXB_GN_MODE=$1 DEXB_GN_LAG=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile > gpurun_out/r02e_m$1_l$2.json 2> gpurun_out/r02e_m$1_l$2_breakdown.txt
  python - <<PY
import json
d=json.load(open("gpurun_out/r02e_m$1_l$2.json"))
print("mode=$1 lag=$2: ms/traj", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["per_bin_violation"])
PY
  grep "gemm+gn" gpurun_out/r02e_m$1_l$2_breakdown.txt
done
