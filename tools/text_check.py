"""One-shot GPU check of the CUDA text encoder against the reference fixtures (tests/golden/text_*.npz):
python tools/text_check.py  -> one line per fixture with max |err| / RMS of mu, logw and of the residual stream after the prenet,
layer 0 and the last layer."""
import glob
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "dex-tts_b200"), os.path.join(ROOT, "tests")]
from dexb200.model import GeTextEncoder, TextEncoder           # noqa: E402
from dexb200.synth import synth_text, synth_text_weights       # noqa: E402
from parity import tensor_rel_err                              # noqa: E402

KW = dict(n_vocab=149, n_feats=80, n_channels=192, filter_channels=1024, filter_channels_dp=256, n_heads=2, n_layers=8, kernel_size=3,
          p_dropout=0.1, use_softmax=True, use_decay=False, window_size=4)
for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "text_*.npz"))):
    if "_spk_" in os.path.basename(path):                       # oracle-only fixture (n_spks > 1)
        continue
    g = np.load(path)
    B, Tx, ragged, seed, dex = [int(v) for v in g["meta"][:5]]
    inp = synth_text(B, Tx, seed=seed, ragged=bool(ragged))
    enc = (TextEncoder if dex else GeTextEncoder)(**KW)
    enc.load_state_dict(synth_text_weights(prefix="", adaln=bool(dex)), strict=True)
    enc = enc.cuda().eval()
    x, xl, sty = inp["x"].cuda(), inp["x_lengths"].cuda(), inp["sty"].cuda()
    try:
        t0 = time.perf_counter()
        mu, logw, x_mask = enc(x, xl, sty) if dex else enc(x, xl)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        mu, logw, x_mask = enc(x, xl, sty) if dex else enc(x, xl)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        eng = enc.cuda_engine()
        e = {"mu": tensor_rel_err(mu.cpu(), torch.from_numpy(g["mu"])), "logw": tensor_rel_err(logw.cpu(), torch.from_numpy(g["logw"]))}
        s = sty if dex else None
        e["prenet"] = tensor_rel_err(eng.forward_stream(x, x_mask, s, 0).cpu(), torch.from_numpy(g["prenet"]).transpose(1, 2))
        e["layer0"] = tensor_rel_err(eng.forward_stream(x, x_mask, s, 1).cpu(), torch.from_numpy(g["layer0"]))
        e["layer7"] = tensor_rel_err(eng.forward_stream(x, x_mask, s, 8).cpu(), torch.from_numpy(g["layer7"]))
        print(os.path.basename(path), " ".join(f"{k} {v:.2e}" for k, v in e.items()), f"launches {eng.launches} first {1e3*(t1-t0):.1f} ms "
              f"second {1e3*(t2-t1):.2f} ms", flush=True)
    except Exception as ex:                                     # keep going: the other fixtures still tell something
        print(os.path.basename(path), "FAILED:", repr(ex)[:300], flush=True)
