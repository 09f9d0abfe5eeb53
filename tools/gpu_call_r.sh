#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipeline --profile > gpurun_out/r02x.json 2> gpurun_out/r02x_breakdown.txt
head -32 gpurun_out/r02x_breakdown.txt
