#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_decoder_gpu.py -x -q 2>&1 | tail -2
for pm in 0 1 0 1; do
  DEXB_PAIR=$pm timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipeline --profile > gpurun_out/r02r_p$pm.json 2> gpurun_out/r02r_p${pm}_breakdown.txt
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02r_p$pm.json").read().strip().splitlines()[-1])
print("pair=$pm: ms/traj", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["per_bin_violation"], "frac", round(d["roofline"]["frac"],4))
PY
  grep -E "conv " gpurun_out/r02r_p${pm}_breakdown.txt
done
timeout 600 ncu --set full --clock-control none -k regex:"conv_pair_kernel|gemm_tc_kernel" -s 60 -c 8 -o /tmp/pairnet python tools/prof_net_call.py C2 2 > gpurun_out/r02r_ncunet.log 2>&1
ncu -i /tmp/pairnet.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/r02r_ncu_pairnet.csv
