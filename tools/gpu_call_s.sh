#!/bin/bash
# warm-cache DRAM traffic per kernel of one network call (single-pass metrics, no cache flush between kernels)
mkdir -p gpurun_out
timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct -s 100 -c 110 --csv --log-file gpurun_out/r02s_warm_dram.csv python tools/prof_net_call.py C2 3 > gpurun_out/r02s.log 2>&1
tail -3 gpurun_out/r02s.log
