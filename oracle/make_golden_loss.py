"""Generate tests/golden/loss_*.npz: the training branch of the UNMODIFIED reference ``Diffusion.forward(infer=False)`` -> ``EDMLoss``
(DEX-TTS/model/diffusion.py:252-254, DEX-TTS/model/edm.py:22-68) on the seeded weights / inputs of the decoder fixtures, with its two
Gaussian draws (``torch.randn([B,1,1])``, ``torch.randn_like(x0)``) patched to seeded CPU tensors.  Modules in eval mode (no dropout),
mask_ratio = 0.  Run in the build container only:   python oracle/make_golden_loss.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "dex-tts_b200"))

import make_golden as MG                                                # noqa: E402
import ref_loader                                                       # noqa: E402
from dexb200.manifest import DecoderCfg                                 # noqa: E402
from dexb200.synth import synth_decoder_weights, synth_inputs           # noqa: E402

CASES = [("loss_dex_b2r", "dex", 2, 48, 19, 61), ("loss_gedex_b2r", "gedex", 2, 52, 0, 62)]      # name, variant, B, T, Ts, seed


def loss_draws(B, T, seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randn(B, 1, 1, generator=g), torch.randn(B, 80, T, generator=g)


def run_case(name, variant, B, T, Ts, seed):
    cfg = DecoderCfg.make(variant)
    dec_cfg, dit_cfg = MG.ref_cfgs(cfg)
    dec, mod = ref_loader.build_reference_decoder(variant, dec_cfg, dit_cfg)
    w = synth_decoder_weights(cfg, seed=100, live=True)
    sd = dict(w)
    sd.update({k.replace("denoise_fn.", "precond_model.model."): v for k, v in w.items()})
    dec.load_state_dict(sd, strict=True)
    inp = synth_inputs(cfg, B, T, Ts=max(Ts, 1), seed=seed, ragged=True)
    x0 = inp["z"]                                                        # any (B,80,T) tensor serves as the clean mel
    rnd, noise = loss_draws(B, T, seed + 1)
    real_randn, real_like = torch.randn, torch.randn_like
    torch.randn = lambda *a, **k: rnd.clone()
    torch.randn_like = lambda t, **k: noise.clone()
    try:
        with torch.no_grad():
            if variant == "dex":
                loss = dec(x0, inp["mask"], inp["mu"], inp["ref_skips"], inp["ref_lengths"], inp["sty"], inp["sty_lengths"], infer=False)
            else:
                loss = dec(x0, inp["mask"], inp["mu"], infer=False)
    finally:
        torch.randn, torch.randn_like = real_randn, real_like
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, loss=np.array(float(loss), dtype=np.float64), meta=np.array([B, T, Ts, seed], dtype=np.int64),
                        variant=np.array(variant))
    print(f"{name}: loss {float(loss):.6f} -> {os.path.relpath(path, ROOT)}")


if __name__ == "__main__":
    torch.set_num_threads(8)
    for c in CASES:
        run_case(*c)
