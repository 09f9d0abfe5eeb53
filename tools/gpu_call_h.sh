#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_libritts_gpu.py tests/test_decoder_gpu.py -x -q -s -k "libri" 2>&1 | tail -40 > gpurun_out/r02h_pytest.log
cat gpurun_out/r02h_pytest.log
