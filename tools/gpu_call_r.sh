#!/bin/bash
mkdir -p gpurun_out
export DEXB_NO_GRAPH=1
timeout 1200 compute-sanitizer --tool synccheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/r02g_sanitizer_synccheck.log 2>&1
echo "synccheck: exit $? -- $(grep -E 'ERROR SUMMARY' gpurun_out/r02g_sanitizer_synccheck.log | tail -1)"
grep "Device Frame" gpurun_out/r02g_sanitizer_synccheck.log | sed 's/(.*//' | sort | uniq -c
unset DEXB_NO_GRAPH
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_decoder_gpu.py -x -q 2>&1 | tail -2
