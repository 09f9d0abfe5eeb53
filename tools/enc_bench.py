"""Time the CUDA style stage (TIV / TV / LF0 encoders + style fusion) on synthetic reference mels:
python tools/enc_bench.py [B] [T] [iters]  -> one JSON line per component."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dex-tts_b200"))
from dexb200.model import LF0Encoder, TIVEncoder, TVEncoder, style_fusion                      # noqa: E402
from dexb200.synth import (synth_conv_sty_weights, synth_lf0, synth_lf0_weights, synth_ref_mel, synth_tiv_weights,   # noqa: E402
                           synth_tv_weights)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 259
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
tiv = TIVEncoder(c_in=80, c_out=64, num_layer=6, c_h=128)
tiv.load_state_dict(synth_tiv_weights(prefix=""), strict=True)
tv = TVEncoder(c_in=80, c_out=192, c_out_g=192, num_layer=6, c_h=128, n_emb=512, commit_w=0.25)
tv.load_state_dict(synth_tv_weights(prefix=""), strict=True)
lf0e = LF0Encoder(c_h=192, c_out=192, c_out_g=192, num_layer=2, c_in=1)
lf0e.load_state_dict(synth_lf0_weights(prefix=""), strict=True)
conv_sty = torch.nn.Conv1d(192, 128, 1, 1)
cw = synth_conv_sty_weights()
conv_sty.load_state_dict({"weight": cw["conv_sty.weight"], "bias": cw["conv_sty.bias"]})
tiv, tv, lf0e, conv_sty = tiv.cuda().eval(), tv.cuda().eval(), lf0e.cuda().eval(), conv_sty.cuda().eval()
inp = synth_ref_mel(B, T, seed=1, ragged=False)
ref, mask = inp["ref"].cuda(), inp["mask"].cuda()
lf0 = synth_lf0(B, T, seed=2)["lf0"].cuda()


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


zb, zd, _ = tv(ref, mask)
le, ld = lf0e(lf0, mask)


def whole():
    tiv(ref, mask)
    z_before, z_dec, _ = tv(ref, mask)
    e, d = lf0e(lf0, mask)
    style_fusion(conv_sty, z_before, z_dec, mask, e, d, mask, want_sty_enc=False)


rows = [("TIV encoder (14 conv1d + BN/ReLU/InstanceNorm1D fusions)", lambda: tiv(ref, mask), tiv.cuda_engine()),
        ("TV encoder (18 conv1d + LayerNorm fusions + fp32 VQ search)", lambda: tv(ref, mask), tv.cuda_engine()),
        ("LF0 encoder (in_conv, 2-layer BiGRU, out_conv, Projection)", lambda: lf0e(lf0, mask), lf0e.cuda_engine()),
        ("style fusion (time mean + conv_sty)", lambda: style_fusion(conv_sty, zb, zd, mask, le, ld, mask, want_sty_enc=False), None),
        ("whole style stage", whole, None)]
for what, fn, eng in rows:
    ms = timed(fn)
    print(json.dumps({"what": what, "B": B, "T": T, "ms": round(ms, 4), "launches": eng.launches if eng is not None else None,
                      "utterances_per_s": round(B / ms * 1e3, 1)}))
