#!/bin/bash
# First GPU call of a round, in priority order (everything lands in gpurun_out/; copy what should be judged to profiles/):
#   gpurun --timeout 900 -- 'bash tools/gpu_first_call.sh r02'
# 1. the whole GPU suite (the run_last items -- text encoder, whole model -- have not been run as pytest on a B200 yet)
# 2. stage checks / micro-benchmarks of the text side and the style stage
# 3. bench lines: default (C2), reference arm, C3 (STFT + style stage inside the step)
# 4. ncu launch list of the bench command (shares of the step, not absolutes)
set -u
TAG="${1:-rXX}"
OUT=gpurun_out
mkdir -p "$OUT"
python -m pytest tests -q -m gpu -s 2>&1 | tail -150 > "$OUT/${TAG}_pytest_gpu.log"
tail -3 "$OUT/${TAG}_pytest_gpu.log"
python tools/text_check.py > "$OUT/${TAG}_text_check.log" 2>&1
python tools/align_bench.py > "$OUT/${TAG}_align_bench.json" 2> "$OUT/${TAG}_align_bench.err"
python tools/tts_bench.py > "$OUT/${TAG}_tts_bench.json" 2> "$OUT/${TAG}_tts_bench.err"
python tools/enc_bench.py > "$OUT/${TAG}_enc_bench.json" 2> "$OUT/${TAG}_enc_bench.err"
python bench.py --steps 5 --warmup 3 > "$OUT/${TAG}_bench_default.json" 2> "$OUT/${TAG}_bench_default.err"
python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/${TAG}_bench_reference.json" 2> "$OUT/${TAG}_bench_reference.err"
python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu-baseline > "$OUT/${TAG}_bench_C3.json" 2> "$OUT/${TAG}_bench_C3.err"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/${TAG}_launches_bench.csv" \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > "$OUT/${TAG}_bench_under_ncu.log" 2>&1     # never a bench value
cat "$OUT/${TAG}_bench_default.json"
