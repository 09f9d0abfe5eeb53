"""Micro-benchmark of the tcgen05 implicit-GEMM engine on the shapes of the C2 workload (tuning aid).
usage: python tools/gemm_bench.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dex-tts_b200"))
import torch  # noqa: E402,F401

from dexb200 import lib as _lib  # noqa: E402

L = _lib.load()
torch.cuda.init()
torch.zeros(1).cuda()
SHAPES = {
    # name: (nimg, H, W, K, N, KH, KW, offH, offW, stride, out_mode)
    "conv L0 64->64": (8, 80, 512, 64, 64, 3, 3, -1, -1, 1, 0),
    "conv L1 128->128": (8, 40, 256, 128, 128, 3, 3, -1, -1, 1, 0),
    "conv L1 256->64": (8, 40, 256, 256, 64, 3, 3, -1, -1, 1, 0),
    "la.kv L0 64->256": (8, 80, 512, 64, 256, 1, 1, 0, 0, 1, 0),
    "qkv 256->768": (8, 1, 2580, 256, 768, 1, 1, 0, 0, 1, 1),
    "fc1 256->512": (1, 1, 20640, 256, 512, 1, 1, 0, 0, 1, 1),
    "fc1 + GELU": (1, 1, 20640, 256, 512, 1, 1, 0, 0, 1, 3),
    "la.vt L0 64->128": (8, 80, 512, 64, 128, 1, 1, 0, 0, 1, 1),
    "fc2 512->256": (1, 1, 20640, 512, 256, 1, 1, 0, 0, 1, 0),
    "lin K=4096 N=32": (1, 1, 37888, 4096, 32, 1, 1, 0, 0, 1, 0),
    "lin K=4096 N=64": (1, 1, 37888, 4096, 64, 1, 1, 0, 0, 1, 0),
    "lin K=4096 N=128": (1, 1, 37888, 4096, 128, 1, 1, 0, 0, 1, 0),
}
for name, (nimg, H, W, K, N, KH, KW, oh, ow, st, om) in SHAPES.items():
    flop = 2.0 * nimg * ((H + st - 1) // st) * ((W + st - 1) // st) * N * K * KH * KW
    row = []
    for dbg in (0, 8, 1, 2, 3, 5):
        ms = ctypes.c_float(0)
        _lib.check(L.dexb_gemm_bench(3, nimg, H, W, K, N, KH, KW, oh, ow, st, om, dbg, 20, ctypes.byref(ms)), "gemm_bench")
        row.append(ms.value)
    print(f"{name:20s} full {row[0]*1e3:7.1f} us ({flop/row[0]*1e-9:6.1f} TFLOP/s)  no-stores {row[1]*1e3:7.1f}  no-epilogue {row[2]*1e3:7.1f}  "
          f"no-mma {row[3]*1e3:7.1f}  neither {row[4]*1e3:7.1f}  mma-only(no TMA, no epilogue) {row[5]*1e3:7.1f}")
