#!/bin/bash
# fused GroupNorm-apply: parity subset, then A/B bench (DEXB_GN_FUSE=0/1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_scale_gpu.py tests/test_taps_gpu.py tests/test_module_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/r02d_pytest.log
cat gpurun_out/r02d_pytest.log
for f in 1 0; do
  DEXB_GN_FUSE=$f timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile > gpurun_out/r02d_bench_f$f.json 2> gpurun_out/r02d_bench_f${f}_breakdown.txt
  python - <<PY
import json
d=json.load(open("gpurun_out/r02d_bench_f$f.json"))
print("fuse=$f: ms/traj", round(d["ms_per_step"],2), "gemm frac", round(d["roofline"]["frac"],4), "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["per_bin_violation"])
PY
done
head -12 gpurun_out/r02d_bench_f1_breakdown.txt
