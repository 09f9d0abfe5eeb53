#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_decoder_gpu.py tests/test_scale_gpu.py tests/test_taps_gpu.py tests/test_libritts_gpu.py -x -q 2>&1 | tail -2
DEXB_BENCH_RANDOM=1 PAIR_BENCH_OUT_MODE=4 timeout 120 python tools/pair_bench.py 2>&1
for tag in fa2e987 HEAD fa2e987 HEAD; do
  if [ $tag = HEAD ]; then dir=.; else dir=ab/$tag; fi
  (cd $dir && timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipeline 2>/dev/null) | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag: ms/traj', round(d['ms_per_step'],2), 'clk', d['clocks']['sm_mhz'], 'parity', d['parity']['per_bin_violation'], 'frac', round(d['roofline']['frac'],4))"
done
