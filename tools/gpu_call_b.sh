#!/bin/bash
# A/B of the MMA-issue-loop changes: grouped halo rounds, early wait
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_decoder_gpu.py tests/test_scale_gpu.py -x -q 2>&1 | tail -5 > gpurun_out/r02b_pytest.log
cat gpurun_out/r02b_pytest.log
for cfg in "1 1" "0 1" "0 0" "1 0"; do
  set -- $cfg
  DEXB_HALO_GROUP=$1 DEXB_EARLY_WAIT=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile > gpurun_out/r02b_bench_g$1_e$2.json 2> gpurun_out/r02b_bench_g$1_e$2_breakdown.txt
  python - <<PY
import json
d=json.load(open("gpurun_out/r02b_bench_g$1_e$2.json"))
print("group=$1 early=$2: ms/traj", round(d["ms_per_step"],2), "gemm frac", round(d["roofline"]["frac"],4), "clk", d["clocks"]["sm_mhz"])
PY
done
