#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_gpu.py tests/test_scale_gpu.py tests/test_libritts_gpu.py -x -q 2>&1 | tail -2
for i in 1 2; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipeline --profile > gpurun_out/r02t.json 2> gpurun_out/r02t_breakdown.txt
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02t.json").read().strip().splitlines()[-1])
print("ms/traj", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["per_bin_violation"], "frac", round(d["roofline"]["frac"],4))
PY
  grep -E "ln_mod|tok_assemble" gpurun_out/r02t_breakdown.txt
done
