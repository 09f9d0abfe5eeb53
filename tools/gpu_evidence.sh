#!/bin/bash
# Evidence run of a round: whole GPU suite, bench lines of every BASELINE config, stage micro-benchmarks, ncu launch list + full
# sections of the tensor-core kernels.  Everything lands in gpurun_out/<TAG>_*; copy what should be judged to profiles/.
#   gpurun --timeout 1800 -- 'bash tools/gpu_evidence.sh r02'
set -u
TAG="${1:-rXX}"
OUT=gpurun_out
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/${TAG}_smi.txt" 2>&1
timeout 1500 python -m pytest tests -q -m gpu -s 2>&1 | tail -220 > "$OUT/${TAG}_pytest_gpu.log"
tail -3 "$OUT/${TAG}_pytest_gpu.log"
timeout 300 python bench.py --steps 5 --warmup 3 --profile > "$OUT/${TAG}_bench_default.json" 2> "$OUT/${TAG}_bench_default_breakdown.txt"
for w in C1 C3 C4 C5; do
  timeout 600 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --profile > "$OUT/${TAG}_bench_$w.json" 2> "$OUT/${TAG}_bench_${w}_breakdown.txt"
done
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > "$OUT/${TAG}_bench_reference.json" 2> "$OUT/${TAG}_bench_reference.err"
timeout 120 python tools/stft_bench.py > "$OUT/${TAG}_stft_bench.json" 2> "$OUT/${TAG}_stft_bench.err"
timeout 120 python tools/enc_bench.py > "$OUT/${TAG}_enc_bench.json" 2> "$OUT/${TAG}_enc_bench.err"
timeout 120 python tools/tts_bench.py > "$OUT/${TAG}_tts_bench.json" 2> "$OUT/${TAG}_tts_bench.err"
timeout 120 python tools/voc_bench.py 8 512 > "$OUT/${TAG}_voc_bench.json" 2> "$OUT/${TAG}_voc_bench.err"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/${TAG}_launches_bench.csv" \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > "$OUT/${TAG}_bench_under_ncu.log" 2>&1     # never a bench value
# the .ncu-rep stays on the box (100 MB: gpurun_out is capped at 64 MiB); only the condensed summary comes back
timeout 900 ncu --set full --clock-control none -k regex:"gemm_tc_kernel|conv_pair_kernel|attn_fwd_kernel|posconv_kernel|k_gn_apply" \
    -s 60 -c 62 -o /tmp/${TAG}_ncu_full python tools/prof_net_call.py C2 2 > "$OUT/${TAG}_ncu_full.log" 2>&1
ncu -i /tmp/${TAG}_ncu_full.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > "$OUT/${TAG}_ncu_summary.csv" 2>> "$OUT/${TAG}_ncu_full.log"
du -sh "$OUT"
ls -la "$OUT" | tail -30
for f in default C1 C3 C4 C5; do python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench_$f.json"))
    print("$f", round(d["ms_per_step"],2), "ms", round(d["value"]), d["unit"], "e2e", round(d["e2e"]["value"]), "parity", d.get("parity",{}).get("per_bin_violation"), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("$f", "failed", e)
PY
done
