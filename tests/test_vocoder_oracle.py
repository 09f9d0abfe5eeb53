"""Pin the oracle restatement of the HiFi-GAN v1 generator (oracle/vocoder_oracle.py; DEX-TTS/hifigan/models.py:96-173) against outputs
of the unmodified reference Generator (tests/golden/voc_*.npz, made by oracle/make_golden_vocoder.py in the build container).
SURVEY.md §8f rank 3; the CUDA side is tests/test_vocoder_gpu.py."""
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


def test_dropin_generator_state_dict():
    """The CUDA drop-in ``hifigan.Generator`` carries the reference generator's parameter tree: weight-normed keys before
    ``remove_weight_norm`` (bias, weight_g, weight_v per convolution, upstream order), the oracle's manifest after it."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), "dex-tts_b200"))
    from dexb200.hifigan import AttrDict, Generator
    cfg = dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
               resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3)
    g = Generator(AttrDict(cfg))
    keys = list(g.state_dict().keys())
    assert len(keys) == 3 * len(V.vocoder_manifest()) // 2
    assert keys[:3] == ["conv_pre.bias", "conv_pre.weight_g", "conv_pre.weight_v"]
    assert tuple(g.state_dict()["ups.0.weight_g"].shape) == (512, 1, 1)          # ConvTranspose1d: norm over dim 0 = input channels
    w = V.synth_vocoder_weights()
    sd = {}
    for name, t in w.items():
        if name.endswith(".bias"):
            sd[name] = t
        else:
            sd[name + "_v"] = t
            sd[name + "_g"] = torch.norm_except_dim(t, 2, 0)
    g.load_state_dict(sd, strict=True)
    g.eval()
    g.remove_weight_norm()
    out = g.state_dict()
    assert {k: tuple(v.shape) for k, v in out.items()} == {k: tuple(s) for k, s in V.vocoder_manifest()}
    for k, v in w.items():
        assert torch.allclose(out[k], v, rtol=1e-6, atol=1e-8), k
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            g(torch.zeros(1, 80, 4))
