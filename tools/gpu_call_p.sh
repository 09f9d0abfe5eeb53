#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_gpu.py tests/test_scale_gpu.py -x -q 2>&1 | tail -2
for w in 1 0 1 0; do
  DEXB_BN_WAVES=$w timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipeline --profile > gpurun_out/r02p_w$w.json 2> gpurun_out/r02p_w${w}_breakdown.txt
  python - <<PY
import json
d=json.load(open("gpurun_out/r02p_w$w.json"))
print("waves=$w: ms/traj", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["per_bin_violation"], "gemm frac", round(d["roofline"]["frac"],4))
PY
  grep -E "k.proj|k.fc1|k.fc2|k.qkv" gpurun_out/r02p_w${w}_breakdown.txt
done
