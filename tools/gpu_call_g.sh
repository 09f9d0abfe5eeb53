#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vocoder_gpu.py tests/test_gemm_gpu.py -x -q -s 2>&1 | tail -25 > gpurun_out/r02g_pytest.log
cat gpurun_out/r02g_pytest.log
timeout 300 python tools/voc_bench.py 8 512 > gpurun_out/r02g_voc_bench.json 2>&1; cat gpurun_out/r02g_voc_bench.json
timeout 300 python tools/voc_bench.py 1 512 >> gpurun_out/r02g_voc_bench.json 2>&1; tail -1 gpurun_out/r02g_voc_bench.json
