#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attn_gpu.py tests/test_decoder_gpu.py tests/test_scale_gpu.py tests/test_tv_gpu.py tests/test_taps_gpu.py -x -q -s 2>&1 | grep -E "attn|passed|failed|Error|error" | tail -30
DEXB_ATTN_ONEPASS=0 timeout 300 python -m pytest tests/test_attn_gpu.py -x -q -s 2>&1 | grep -E "attn|passed|failed" | tail -14
for o in 1 0 1 0; do
  DEXB_ATTN_ONEPASS=$o timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile > gpurun_out/r02i_o$o.json 2> gpurun_out/r02i_o${o}_breakdown.txt
  python - <<PY
import json
d=json.load(open("gpurun_out/r02i_o$o.json"))
print("onepass=$o: ms/traj", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["per_bin_violation"])
PY
  grep "attn_fwd" gpurun_out/r02i_o${o}_breakdown.txt
done
for o in 1 0; do
  DEXB_ATTN_ONEPASS=$o timeout 300 python bench.py --workload C5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_C5_o$o.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/r02i_C5_o$o.json"))
print("C5 onepass=$o: ms/traj", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["per_bin_violation"])
PY
done
