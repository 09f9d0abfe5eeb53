"""Convolution shapes of the C2 workload through dexb_gemm_bench (full kernel): run once with DEXB_PAIR=0 and once with DEXB_PAIR=1.
usage: DEXB_PAIR=1 python tools/pair_bench.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dex-tts_b200"))
import torch  # noqa: E402,F401

from dexb200 import lib as _lib  # noqa: E402

L = _lib.load()
OUT_MODE = int(os.environ.get('PAIR_BENCH_OUT_MODE', '0'))       # 4 = with GroupNorm sums in the epilogue
torch.zeros(1).cuda()
SHAPES = {
    "conv L0 64->64": (8, 80, 512, 64, 64),
    "conv L0 128->64": (8, 80, 512, 128, 64),
    "conv L1 128->128": (8, 40, 256, 128, 128),
    "conv L1 256->128": (8, 40, 256, 256, 128),
    "conv L2 256->256": (8, 20, 128, 256, 256),
    "conv L0 64->64 B32": (32, 80, 512, 64, 64),
    "conv L1 128->128 B32": (32, 40, 256, 128, 128),
}
for name, (nimg, H, W, K, N) in SHAPES.items():
    flop = 2.0 * nimg * H * W * N * K * 9
    best = 1e9
    for _ in range(3):
        ms = ctypes.c_float(0)
        _lib.check(L.dexb_gemm_bench(3, nimg, H, W, K, N, 3, 3, -1, -1, 1, OUT_MODE, 0, 20, ctypes.byref(ms)), "gemm_bench")
        best = min(best, ms.value)
    print(f"DEXB_PAIR={os.environ.get('DEXB_PAIR', '1')} out_mode={OUT_MODE}  {name:24s} {best*1e3:7.1f} us ({flop/best*1e-9:6.1f} TFLOP/s)")
