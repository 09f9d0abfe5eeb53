#!/bin/bash
# r02 re-entry: whole GPU suite + C2 bench line with the per-kernel breakdown + reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02c_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -150 > gpurun_out/r02c_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --profile > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench_breakdown.txt
tail -3 gpurun_out/r02c_pytest_gpu.log
cat gpurun_out/r02c_bench.json
