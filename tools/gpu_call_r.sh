#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02v.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r02v.json").read().strip().splitlines()[-1])
print("ms/traj", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], "e2e", round(d["e2e"]["value"]), "pipeline", d.get("pipeline",{}).get("vocoder_ms"))
PY
timeout 300 python tools/enc_bench.py 2>&1 | tail -6
timeout 300 python tools/tts_bench.py 2>&1 | tail -2
timeout 300 python tools/voc_bench.py 2>&1 | tail -2
