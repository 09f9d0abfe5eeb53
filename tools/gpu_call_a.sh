#!/bin/bash
# r02 first GPU call: whole GPU suite (incl. the new at-scale parity + tap tests), issue-loop micro-benchmark, C2 bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -120 > gpurun_out/r02a_pytest_gpu.log
timeout 120 tools/_bin/issue_bench > gpurun_out/r02a_issue_bench.txt 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --profile > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench_breakdown.txt
tail -3 gpurun_out/r02a_pytest_gpu.log
cat gpurun_out/r02a_issue_bench.txt
