"""Generate tests/golden/lf0_*.npz by running the UNMODIFIED reference LF0Encoder (DEX-TTS/model/ref_encoder.py:36-56) and the
style-fusion lines of DeXTTS.forward (DEX-TTS/model/tts.py:45-49, executed verbatim on the reference modules' outputs with a real
``nn.Conv1d`` as ``conv_sty``).  Run in the build container only:   python oracle/make_golden_lf0.py
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "dex-tts_b200"))

import ref_loader                                     # noqa: E402
from dexb200.synth import synth_conv_sty_weights, synth_lf0, synth_lf0_weights, synth_ref_mel, synth_tv_weights   # noqa: E402

LF0_CFG = dict(c_in=1, c_h=192, c_out=192, c_out_g=192, num_layer=2)                 # DEX-TTS/config/VCTK/base.yaml:37-42
TV_CFG = dict(c_in=80, num_layer=6, c_h=128, c_out=192, c_out_g=192, commit_w=0.25, n_emb=512)
CASES = [
    # name,       B, T,   ragged, seed
    ("lf0_b1",    1, 41,  False, 51),
    ("lf0_b2r",   2, 150, True,  52),
]


def run_case(name, B, T, ragged, seed):
    ref_loader.load_reference("dex")
    enc_mod = importlib.import_module("model.ref_encoder")
    enc = enc_mod.LF0Encoder(**LF0_CFG)
    enc.load_state_dict(synth_lf0_weights(**LF0_CFG, seed=100, prefix=""), strict=True)
    enc.eval()
    tv = enc_mod.TVEncoder(**TV_CFG)
    tv.load_state_dict(synth_tv_weights(**{k: v for k, v in TV_CFG.items() if k != "commit_w"}, seed=100, prefix=""), strict=True)
    tv.eval()
    conv_sty = torch.nn.Conv1d(192, 128, 1, 1)                                       # tts.py:31
    cw = synth_conv_sty_weights()
    conv_sty.load_state_dict({"weight": cw["conv_sty.weight"], "bias": cw["conv_sty.bias"]})
    inp = synth_lf0(B, T, seed=seed, ragged=ragged)
    sty_in = synth_ref_mel(B, T, seed=seed + 100, ragged=ragged)                     # style mel of the same length (synthesize.py)
    with torch.no_grad():
        lf0_mask, sty_mask = inp["mask"], sty_in["mask"]
        lf0_enc, lf0_dec = enc(inp["lf0"], lf0_mask)                                 # tts.py:42
        sty_enc, sty_dec, _ = tv(sty_in["ref"].unsqueeze(1), sty_mask)               # tts.py:43
        # tts.py:45-49, verbatim
        sty_enc = (sty_enc.sum(dim=-1) / sty_mask.sum(dim=-1)) + (lf0_enc.sum(dim=-1) / lf0_mask.sum(dim=-1))
        sty_enc = sty_enc.squeeze(1)
        sty_dec = sty_dec + (lf0_dec.sum(dim=-1) / lf0_mask.sum(dim=-1)).unsqueeze(-1)
        sty_dec = conv_sty(sty_dec)
    arrs = dict(lf0_enc=lf0_enc.numpy(), lf0_dec=lf0_dec.numpy(), sty_enc=sty_enc.numpy(), sty_dec=sty_dec.numpy(),
                meta=np.array([B, T, int(ragged), seed], dtype=np.int64), keys=np.array(list(enc.state_dict().keys())))
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: lf0_enc {tuple(lf0_enc.shape)} |max| {float(lf0_enc.abs().max()):.3f} lf0_dec |max| {float(lf0_dec.abs().max()):.3f} "
          f"sty_dec {tuple(sty_dec.shape)} |max| {float(sty_dec.abs().max()):.3f} -> {os.path.relpath(path, ROOT)} "
          f"({os.path.getsize(path)/1024:.0f} KiB)")


if __name__ == "__main__":
    torch.set_num_threads(8)
    for c in CASES:
        run_case(*c)
