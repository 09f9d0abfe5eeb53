#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_scale_gpu.py tests/test_taps_gpu.py tests/test_module_gpu.py -x -q 2>&1 | tail -3
for i in 1 2; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipeline --profile > gpurun_out/r02j_$i.json 2> gpurun_out/r02j_${i}_breakdown.txt
  python - <<PY
import json
d=json.load(open("gpurun_out/r02j_$i.json"))
print("run $i: ms/traj", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["per_bin_violation"], "gemm frac", round(d["roofline"]["frac"],4))
PY
  grep -E "gn_apply|conv_in" gpurun_out/r02j_${i}_breakdown.txt
done
