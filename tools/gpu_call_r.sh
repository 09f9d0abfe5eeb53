#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"conv_pair_kernel" -s 14 -c 2 -o /tmp/cp python tools/prof_net_call.py C2 2 > gpurun_out/r02w_ncu.log 2>&1
ncu -i /tmp/cp.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/r02w_ncu_cp.csv
ncu -i /tmp/cp.ncu-rep --page source --csv --print-source sass 2>/dev/null > gpurun_out/r02w_cp_source.csv
ls -la gpurun_out/r02w*
