"""Generate tests/golden/tiv_*.npz by running the UNMODIFIED reference TIVEncoder (DEX-TTS/model/ref_encoder.py:83-107,
imported from /root/reference).  Run in the build container only:   python oracle/make_golden_tiv.py

For every case: build ``TIVEncoder(**cfg.tiv_encoder)`` with the values of DEX-TTS/config/VCTK/base.yaml:44-48, load the seeded
synthetic tensors of ``dexb200.synth.synth_tiv_weights`` with ``load_state_dict(strict=True)`` (pins key / shape compatibility
with upstream checkpoints), ``eval()``, run ``forward(ref, mask)`` exactly as DeXTTS.forward does (tts.py:38,50) and store the
output, the six skip tensors and the state-dict key list.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "dex-tts_b200"))

import importlib                                      # noqa: E402

import ref_loader                                     # noqa: E402
from dexb200.synth import synth_ref_mel, synth_tiv_weights   # noqa: E402

TIV_CFG = dict(c_in=80, num_layer=6, c_h=128, c_out=64)            # DEX-TTS/config/VCTK/base.yaml:44-48
CASES = [
    # name,      B, T,   ragged, seed
    ("tiv_b1",   1, 41,  False, 31),
    ("tiv_b3r",  3, 150, True,  32),       # two 128-frame tiles per utterance, ragged lengths
]


def run_case(name, B, T, ragged, seed):
    ref_loader.load_reference("dex")                  # registers the bare `model` namespace package
    enc_mod = importlib.import_module("model.ref_encoder")
    enc = enc_mod.TIVEncoder(**TIV_CFG)
    w = synth_tiv_weights(**TIV_CFG, seed=100, prefix="")
    enc.load_state_dict(w, strict=True)
    enc.eval()
    inp = synth_ref_mel(B, T, seed=seed, ragged=ragged)
    with torch.no_grad():
        out, skips = enc(inp["ref"].unsqueeze(1), inp["mask"])       # synthesize.py feeds (B,1,80,T); forward squeezes it
    arrs = dict(out=out.numpy(), meta=np.array([B, T, int(ragged), seed], dtype=np.int64),
                keys=np.array(list(enc.state_dict().keys())))
    for i, s_ in enumerate(skips):
        arrs[f"skip{i}"] = s_.numpy()
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: out {tuple(out.shape)} skips {len(skips)} x {tuple(skips[0].shape)} |skip5|max "
          f"{float(skips[-1].abs().max()):.3f} -> {os.path.relpath(path, ROOT)} ({os.path.getsize(path)/1024:.0f} KiB)")


if __name__ == "__main__":
    torch.set_num_threads(8)
    for c in CASES:
        run_case(*c)
