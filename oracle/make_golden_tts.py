"""Generate tests/golden/tts_*.npz by running the UNMODIFIED reference models end to end: ``DeXTTS(cfg.model).forward`` (DEX-TTS/model/
tts.py:33-74) and ``GeDEXTTS(cfg.model).forward`` (GeDEX-TTS/model/tts.py:27-56), built from the reference's own config yaml as
synthesize.py:67 does, with the seeded weights of dexb200.synth loaded ``strict=True`` and ``torch.randn`` patched to a seeded CPU
draw (the reference draws the initial noise on the device; its shape depends on the predicted durations).  Batched DEX-TTS needs sigma
broadcast to (B,) (reference bug), as in make_golden.py.  Run in the build container only:
    python oracle/make_golden_tts.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "dex-tts_b200"))

import ref_loader                                                                    # noqa: E402
from dexb200.synth import reference_state_dict, seeded_noise, synth_lf0, synth_ref_mel, synth_text, synth_tts_weights   # noqa: E402

CASES = [
    # name,          variant, B, Tx, Ts, steps, ragged, seed, temperature, length_scale
    # (text lengths chosen so that the predicted mel lengths land in the range the decoder's own fixtures cover: 44 ... 64 frames)
    ("tts_dex_b1",   "dex",   1, 22, 37, 3,     False,  91,   1.5,         1.0),
    ("tts_dex_b2r",  "dex",   2, 24, 29, 2,     True,   92,   1.5,         1.0),
    ("tts_gedex_b2r", "gedex", 2, 38, 0, 3,     True,   93,   1.5,         1.0),
    # DEX-TTS/config/LibriTTS/base.yaml (decoder dim 128, DiT hidden 384, 256-wide encoders); the name does not match tts_*.npz on
    # purpose: the VCTK-size tests glob that pattern
    ("libritts_dex_b1", "dex", 1, 20, 31, 2,    False,  94,   1.5,         1.0, "LibriTTS"),
    ("libritts_dex_b2r", "dex", 2, 21, 27, 2,   True,   95,   1.5,         1.0, "LibriTTS"),
]


def synth_tts_inputs(variant, B, Tx, Ts, seed, ragged, c_sty=192):
    inp = synth_text(B, Tx, c_sty=c_sty, seed=seed, ragged=ragged)
    if variant == "dex":
        mel = synth_ref_mel(B, Ts, seed=seed + 1, ragged=ragged)             # synthesize.py:94-97: ref = sty = the reference mel
        lf0 = synth_lf0(B, Ts, seed=seed + 2, ragged=ragged)
        inp.update(ref=mel["ref"], ref_lengths=mel["ref_lengths"], lf0=lf0["lf0"], lf0_lengths=mel["ref_lengths"])
    return inp


def run_case(name, variant, B, Tx, Ts, steps, ragged, seed, temperature, length_scale, dataset="VCTK"):
    model, mod, _ = ref_loader.build_reference_tts(variant, dataset=dataset if dataset != "VCTK" else None)
    model.load_state_dict(reference_state_dict(synth_tts_weights(variant, dataset=dataset)), strict=True)
    inp = synth_tts_inputs(variant, B, Tx, Ts, seed, ragged)
    if B > 1 and variant == "dex":
        edm = sys.modules["model.edm"]
        orig = edm.EDMPrecond.forward
        edm.EDMPrecond.forward = lambda self, x, sigma, *a, **k: orig(self, x, sigma.reshape(-1).expand(x.shape[0]), *a, **k)
    cap = {}
    if variant == "dex":                          # the decoder's `sty` input as the reference computed it (tts.py:48-49,71)
        model.decoder.register_forward_pre_hook(lambda m, args: cap.__setitem__("sty_dec", args[5].detach().numpy().copy()))
    noise = seeded_noise(seed + 3)
    real_randn = torch.randn
    torch.randn = lambda *shape, **k: noise(tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape))
    try:
        with torch.no_grad():
            if variant == "dex":
                enc_out, dec_out, attn = model(inp["x"], inp["x_lengths"], inp["ref"], inp["ref_lengths"], inp["ref"], inp["ref_lengths"],
                                               inp["lf0"], inp["lf0_lengths"], n_timesteps=steps, temperature=temperature, spk=None,
                                               length_scale=length_scale)
            else:
                enc_out, dec_out, attn = model(inp["x"], inp["x_lengths"], n_timesteps=steps, temperature=temperature, spk=None,
                                               length_scale=length_scale)
    finally:
        torch.randn = real_randn
    sd = model.state_dict()                       # key / shape list of the reference model: the layout a drop-in has to reproduce
    cap["keys"] = np.array(list(sd.keys()))
    cap["shapes"] = np.array([",".join(str(n) for n in v.shape) for v in sd.values()])
    arrs = dict(enc_out=enc_out.numpy(), dec_out=dec_out.numpy(), attn=np.packbits(attn.numpy().astype(np.uint8), axis=-1),
                attn_shape=np.array(attn.shape, dtype=np.int64), variant=np.array(variant), dataset=np.array(dataset),
                meta=np.array([B, Tx, Ts, steps, int(ragged), seed], dtype=np.int64), scale=np.array([temperature, length_scale]), **cap)
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: enc_out {tuple(enc_out.shape)} dec_out |max| {float(dec_out.abs().max()):.3f} attn {tuple(attn.shape)} -> "
          f"{os.path.relpath(path, ROOT)} ({os.path.getsize(path)/1024:.0f} KiB)")


if __name__ == "__main__":
    torch.set_num_threads(8)
    only = set(sys.argv[1:])
    for c in CASES:
        if not only or c[0] in only:
            run_case(*c)
