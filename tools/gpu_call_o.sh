#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tiv_gpu.py tests/test_tv_gpu.py tests/test_lf0_gpu.py tests/test_text_gpu.py tests/test_tts_gpu.py tests/test_libritts_gpu.py -x -q 2>&1 | tail -4
timeout 120 python tools/enc_bench.py > gpurun_out/r02o_enc_bench.json 2> gpurun_out/r02o_enc_bench.err; cat gpurun_out/r02o_enc_bench.json
timeout 120 python tools/tts_bench.py > gpurun_out/r02o_tts_bench.json 2> gpurun_out/r02o_tts_bench.err; cat gpurun_out/r02o_tts_bench.json
DEXB_NO_GRAPH=1 timeout 120 python tools/tts_bench.py 8 128 259 4 > gpurun_out/r02o_tts_bench_nograph.json 2>/dev/null; cat gpurun_out/r02o_tts_bench_nograph.json
timeout 120 python tools/tts_bench.py 8 128 259 4 > gpurun_out/r02o_tts_bench_graph4.json 2>/dev/null; cat gpurun_out/r02o_tts_bench_graph4.json
