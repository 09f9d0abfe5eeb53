#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 11 -c 1 -o gpurun_out/r02k_attn python tools/prof_net_call.py C2 2 > gpurun_out/r02k_ncu.log 2>&1
ls -la gpurun_out/r02k_attn.ncu-rep
ncu -i gpurun_out/r02k_attn.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/r02k_attn_summary.csv
ncu -i gpurun_out/r02k_attn.ncu-rep --page source --csv > gpurun_out/r02k_attn_source.csv 2>/dev/null
du -sh gpurun_out
