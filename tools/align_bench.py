"""Time the duration / alignment glue (dexb_align_lengths + host round trip + dexb_align_expand) at the text lengths of
BASELINE.json's configs:  python tools/align_bench.py [iters]  -> one JSON line per shape (CUDA events around the whole call,
which contains the path's one stream synchronisation, and around dexb_align_expand alone)."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dex-tts_b200"))
from dexb200 import lib                                       # noqa: E402
from dexb200.model import align_durations                     # noqa: E402
from dexb200.synth import synth_align_inputs                  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50


def timed(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


L = lib.load()
p = lambda t: ctypes.c_void_p(t.data_ptr())
for B, Tx in ((8, 128), (32, 128), (8, 512)):
    inp = synth_align_inputs(B, Tx, seed=1, mean_dur=4.0)
    logw, x_mask, mu_x = inp["logw"].cuda(), inp["x_mask"].cuda(), inp["mu_x"].cuda()
    mu_y, y_mask, attn, y_lengths, y_max = align_durations(logw, x_mask, mu_x)
    Ty = mu_y.shape[-1]
    ms_all = timed(lambda: align_durations(logw, x_mask, mu_x))
    cum = torch.cumsum(torch.ceil(torch.exp(logw) * x_mask).reshape(B, Tx), 1).contiguous()
    xm = x_mask.reshape(B, Tx).contiguous()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ms_exp = timed(lambda: L.dexb_align_expand(p(cum), p(xm), p(y_lengths), p(mu_x), B, Tx, 80, Ty, p(attn), p(y_mask), p(mu_y), st))
    ms_lean = timed(lambda: L.dexb_align_expand(p(cum), p(xm), p(y_lengths), p(mu_x), B, Tx, 80, Ty, None, p(y_mask), p(mu_y), st))
    out_bytes = 4 * B * (Tx * Ty + 80 * Ty + Ty)
    print(json.dumps({"what": "duration / alignment glue", "B": B, "Tx": Tx, "Ty": Ty, "ms_whole_call_incl_host_sync": round(ms_all, 4),
                      "ms_expand": round(ms_exp, 4), "ms_expand_no_attn": round(ms_lean, 4), "expand_out_MB": round(out_bytes / 1e6, 2),
                      "expand_GBps": round(out_bytes / ms_exp / 1e6, 1)}))
