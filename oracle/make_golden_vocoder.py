"""Generate tests/golden/voc_*.npz by running the UNMODIFIED reference HiFi-GAN ``Generator`` (DEX-TTS/hifigan/models.py:112-173, built
from DEX-TTS/hifigan/config.json, put in the state ``get_vocoder`` leaves it in: eval + remove_weight_norm, DEX-TTS/src/utils.py:276-279)
with the seeded weights of oracle/vocoder_oracle.py.  Run in the build container only:   python oracle/make_golden_vocoder.py
"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)

import vocoder_oracle as V                     # noqa: E402

REF = os.path.join(os.environ.get("DEX_REFERENCE_ROOT", "/root/reference"), "DEX-TTS", "hifigan")
CASES = [("voc_b1", 1, 9, 71), ("voc_b2", 2, 14, 72)]          # name, B, mel frames, seed


class AttrDict(dict):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.__dict__ = self


def run_case(name, B, T, seed):
    spec = importlib.util.spec_from_file_location("ref_hifigan_models", os.path.join(REF, "models.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(REF, "config.json")) as f:
        gen = mod.Generator(AttrDict(json.load(f)))
    gen.eval()
    gen.remove_weight_norm()
    w = V.synth_vocoder_weights()
    assert {k: tuple(v.shape) for k, v in gen.state_dict().items()} == dict(V.vocoder_manifest())     # (remove_weight_norm re-orders keys)
    gen.load_state_dict(w, strict=True)
    g = torch.Generator()
    g.manual_seed(seed)
    mel = torch.randn(B, 80, T, generator=g) * 1.5 - 4.0
    with torch.no_grad():
        wav = gen(mel)
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, wav=wav.numpy(), meta=np.array([B, T, seed], dtype=np.int64))
    print(f"{name}: wav {tuple(wav.shape)} |max| {float(wav.abs().max()):.3f} rms {float(wav.pow(2).mean().sqrt()):.3f} -> "
          f"{os.path.relpath(path, ROOT)} ({os.path.getsize(path)/1024:.0f} KiB)")


if __name__ == "__main__":
    torch.set_num_threads(8)
    for c in CASES:
        run_case(*c)
