"""Generate tests/golden/tv_*.npz by running the UNMODIFIED reference TVEncoder (DEX-TTS/model/ref_encoder.py:109-140,
imported from /root/reference).  Run in the build container only:   python oracle/make_golden_tv.py

For every case: build ``TVEncoder(**cfg.tv_encoder)`` with the values of DEX-TTS/config/VCTK/base.yaml:28-35, load the seeded
synthetic tensors of ``dexb200.synth.synth_tv_weights`` with ``load_state_dict(strict=True)`` (pins key / shape compatibility
with upstream checkpoints), ``eval()``, run ``forward(sty, mask)`` exactly as DeXTTS.forward does (tts.py:40,43) and store
z_beforeVQ, z_dec, vq_loss, the code indices the reference picked and the state-dict key list.
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
from dexb200.synth import synth_ref_mel, synth_tv_weights   # noqa: E402

TV_CFG = dict(c_in=80, num_layer=6, c_h=128, c_out=192, c_out_g=192, commit_w=0.25, n_emb=512)   # DEX-TTS/config/VCTK/base.yaml:28-35
CASES = [
    # name,      B, T,   ragged, seed
    ("tv_b1",    1, 41,  False, 41),
    ("tv_b2r",   2, 150, True,  42),       # two 128-frame tiles per utterance, ragged lengths
]


def run_case(name, B, T, ragged, seed):
    ref_loader.load_reference("dex")                  # registers the bare `model` namespace package
    enc_mod = importlib.import_module("model.ref_encoder")
    enc = enc_mod.TVEncoder(**TV_CFG)
    w = synth_tv_weights(**{k: v for k, v in TV_CFG.items() if k != "commit_w"}, seed=100, prefix="")
    enc.load_state_dict(w, strict=True)
    enc.eval()
    inp = synth_ref_mel(B, T, seed=seed, ragged=ragged)
    picked = {}
    real_argmin = torch.argmin

    def spy(*a, **k):
        r = real_argmin(*a, **k)
        picked["idx"] = r.clone()
        return r
    torch.argmin = spy                                              # record the codes VQEmbeddingEMA.forward picks (:211)
    try:
        with torch.no_grad():
            z_before, z_dec, loss = enc(inp["ref"].unsqueeze(1), inp["mask"])   # synthesize.py feeds (B,1,80,T)
    finally:
        torch.argmin = real_argmin
    arrs = dict(z_before=z_before.numpy(), z_dec=z_dec.numpy(), vq_loss=np.array(float(loss), dtype=np.float32),
                idx=picked["idx"].numpy().reshape(B, T), meta=np.array([B, T, int(ragged), seed], dtype=np.int64),
                keys=np.array(list(enc.state_dict().keys())))
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: z_before {tuple(z_before.shape)} |z|max {float(z_before.abs().max()):.3f} z_dec |max| {float(z_dec.abs().max()):.3f} "
          f"loss {float(loss):.4f} distinct codes {len(set(arrs['idx'].ravel().tolist()))} -> {os.path.relpath(path, ROOT)} "
          f"({os.path.getsize(path)/1024:.0f} KiB)")


if __name__ == "__main__":
    torch.set_num_threads(8)
    for c in CASES:
        run_case(*c)
