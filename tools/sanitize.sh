#!/bin/bash
# compute-sanitizer leg: memcheck + racecheck + synccheck over one small reverse diffusion of both variants (the
# smoke test: every kernel class of the loop -- tcgen05 GEMM / attention / pos-conv with their mbarrier + TMEM protocols), the vocoder and
# the STFT.  Slow (minutes): run on demand, not in the pytest suite.   gpurun --timeout 1500 -- 'bash tools/sanitize.sh r02'
set -u
TAG="${1:-rXX}"
OUT=gpurun_out
mkdir -p "$OUT"
export DEXB_NO_GRAPH=1          # plain launches: the sanitizer attributes errors to kernels, not to graph nodes
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 7 python __graft_entry__.py smoke > "$OUT/${TAG}_sanitizer_${tool}.log" 2>&1
  echo "$tool: exit $? -- $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/${TAG}_sanitizer_${tool}.log" | tail -1)"
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_vocoder_gpu.py tests/test_stft_gpu.py -x -q -k "oracle or fixture" > "$OUT/${TAG}_sanitizer_memcheck_voc_stft.log" 2>&1
echo "memcheck vocoder+stft: exit $? -- $(grep -E 'ERROR SUMMARY' "$OUT/${TAG}_sanitizer_memcheck_voc_stft.log" | tail -1)"
