"""Pin the oracle restatement of the HiFi-GAN v1 generator (oracle/vocoder_oracle.py; DEX-TTS/hifigan/models.py:96-173) against outputs
of the unmodified reference Generator (tests/golden/voc_*.npz, made by oracle/make_golden_vocoder.py in the build container).
SURVEY.md §8f rank 3: the oracle of a stage that has no CUDA side yet, so there is no GPU test beside this file."""
import glob
import os

import numpy as np
import pytest
import torch

import vocoder_oracle as V
from parity import tensor_rel_err

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "voc_*.npz")))


def test_golden_present():
    assert len(GOLD) >= 2


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_reference(path):
    g = np.load(path)
    B, T, seed = [int(v) for v in g["meta"]]
    gen = torch.Generator()
    gen.manual_seed(seed)
    mel = torch.randn(B, 80, T, generator=gen) * 1.5 - 4.0
    with torch.no_grad():
        wav = V.hifigan_generator(V.synth_vocoder_weights(), mel)
    assert wav.shape == (B, 1, 256 * T) == g["wav"].shape
    err = tensor_rel_err(wav, torch.from_numpy(g["wav"]))
    rms = float(wav.pow(2).mean().sqrt())
    print(f"vocoder fixture {os.path.basename(path)}: {err:.2e}, rms {rms:.3f}")
    assert err < 2e-5                       # measured 0 (bit-exact) in the build container
    assert 0.05 < rms < 0.9                 # the seeded weights keep tanh out of both dead zones
