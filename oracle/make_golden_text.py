"""Generate tests/golden/text_*.npz by running the UNMODIFIED reference TextEncoder (DEX-TTS/model/text_encoder.py:97-142 over
RetNetModel, DEX-TTS/model/retnet.py / retention.py) on seeded phoneme ids and style vectors with the seeded weights of
dexb200.synth.synth_text_weights.  Run in the build container only:   python oracle/make_golden_text.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "dex-tts_b200"))

import ref_loader                                                     # noqa: E402
from dexb200.synth import synth_text, synth_text_weights              # noqa: E402

ENC_CFG = dict(n_channels=192, filter_channels=1024, filter_channels_dp=256, n_layers=8, kernel_size=3, p_dropout=0.1, n_heads=2,
               window_size=4, use_softmax=True, use_decay=False)     # DEX-TTS/config/VCTK/base.yaml:51-61
CASES = [
    # name,            variant, B, Tx,  ragged, seed
    ("text_b1",        "dex",   1, 19,  False,  81),
    ("text_b2r",       "dex",   2, 128, True,   82),      # the phoneme length of BASELINE.json's C2 / C3, one padded utterance
    ("text_gedex_b2r", "gedex", 2, 45,  True,   83),      # GeDEX-TTS: the same encoder without AdaLN / style (same yaml values)
    ("text_gedex_spk_b2r", "gedex", 2, 33, True, 84, 4),  # n_spks = 4: speaker embedding concatenated behind the prenet (256 wide); oracle only
]


def run_case(name, variant, B, Tx, ragged, seed, n_spks=1):
    spk_dim = 64 if n_spks > 1 else 0
    enc, _ = ref_loader.build_reference_text_encoder(ENC_CFG, variant=variant, n_spks=n_spks, spk_emb_dim=64)
    sd = synth_text_weights(prefix="", adaln=variant == "dex", spk_emb_dim=spk_dim)
    assert list(sd.keys()) == list(enc.state_dict().keys()), "manifest order differs from the reference state_dict"
    enc.load_state_dict(sd, strict=True)
    inp = synth_text(B, Tx, seed=seed, ragged=ragged)
    taps = {}
    hooks = [enc.prenet.register_forward_hook(lambda m, i, o: taps.__setitem__("prenet", o.detach().numpy().copy()))]
    for l in (0, 7):
        hooks.append(enc.encoder.layers[l].register_forward_hook(
            lambda m, i, o, l=l: taps.__setitem__(f"layer{l}", o[0].detach().numpy().copy())))
    with torch.no_grad():
        spk = torch.randn(B, 64, generator=torch.Generator().manual_seed(seed + 7)) if n_spks > 1 else None   # spk_emb(spk), G tts.py:30-31
        mu, logw, x_mask = enc(inp["x"], inp["x_lengths"], inp["sty"]) if variant == "dex" else enc(inp["x"], inp["x_lengths"], spk=spk)
    for h in hooks:
        h.remove()
    arrs = dict(mu=mu.numpy(), logw=logw.numpy(), x_mask=x_mask.numpy(), meta=np.array([B, Tx, int(ragged), seed, int(variant == "dex"), n_spks], dtype=np.int64),
                keys=np.array(list(enc.state_dict().keys())), **taps)
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: mu {tuple(mu.shape)} |max| {float(mu.abs().max()):.3f} logw |max| {float(logw.abs().max()):.3f} -> "
          f"{os.path.relpath(path, ROOT)} ({os.path.getsize(path)/1024:.0f} KiB)")


if __name__ == "__main__":
    torch.set_num_threads(8)
    for c in CASES:
        run_case(*c)
