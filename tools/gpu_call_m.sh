#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile > gpurun_out/r02m.json 2> gpurun_out/r02m_breakdown.txt
python - <<PY
import json
d=json.load(open("gpurun_out/r02m.json"))
print("ms/traj", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["per_bin_violation"], "gemm frac", round(d["roofline"]["frac"],4))
PY
