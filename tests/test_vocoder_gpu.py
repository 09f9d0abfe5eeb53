"""CUDA HiFi-GAN v1 generator (csrc/vocoder.cu) through the C ABI (``dexb_voc_*``) and the drop-in
``hifigan.Generator``: against outputs of the UNMODIFIED reference Generator (tests/golden/voc_*.npz) and against the CPU oracle
(oracle/vocoder_oracle.py, bit-identical to the reference on those fixtures) at the lengths the loop is benchmarked at.
Tolerance: the path's 1e-3, relative to the waveform's RMS (tests/parity.py::tensor_rel_err)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

import vocoder_oracle as V
from parity import tensor_rel_err

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "voc_*.npz")))
CFG = dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
           resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]])   # DEX-TTS/hifigan/config.json


@pytest.fixture(scope="module")
def engine():
    from dexb200.hifigan.models import VocoderEngine
    eng = VocoderEngine(CFG)
    eng.load_state_dict(V.synth_vocoder_weights())
    yield eng
    eng.close()


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_matches_reference_fixture(engine, path):
    g = np.load(path)
    B, T, seed = [int(v) for v in g["meta"]]
    gen = torch.Generator()
    gen.manual_seed(seed)
    mel = torch.randn(B, 80, T, generator=gen) * 1.5 - 4.0
    wav = engine.forward(mel.cuda()).cpu()
    assert wav.shape == g["wav"].shape
    err = tensor_rel_err(wav, torch.from_numpy(g["wav"]))
    print(f"vocoder fixture {os.path.basename(path)}: {err:.2e}, launches {engine.launches}")
    assert err < 1e-3
    assert float(wav.abs().max()) <= 1.0


@pytest.mark.parametrize("B,T", [(1, 1), (2, 37), (1, 200), (3, 130)])
def test_matches_oracle(engine, B, T):
    """ragged tile counts: T * 8 / 64 / 128 / 256 samples are not multiples of the 128-row tile, and B > 1 exercises the zero padding
    at both ends of every utterance (rows of neighbouring utterances are adjacent in memory)"""
    gen = torch.Generator()
    gen.manual_seed(500 + 7 * B + T)
    mel = torch.randn(B, 80, T, generator=gen) * 1.5 - 4.0
    with torch.no_grad():
        ref = V.hifigan_generator(V.synth_vocoder_weights(), mel)
    wav = engine.forward(mel.cuda()).cpu()
    err = tensor_rel_err(wav, ref)
    print(f"vocoder B={B} T={T}: {err:.2e}")
    assert err < 1e-3


def test_benchmark_length_first_samples(engine):
    """T = 512 mel frames (the loop's benchmark length, 131 072 samples per utterance): the CPU oracle on one of two utterances."""
    gen = torch.Generator()
    gen.manual_seed(77)
    mel = torch.randn(2, 80, 512, generator=gen) * 1.5 - 4.0
    with torch.no_grad():
        ref = V.hifigan_generator(V.synth_vocoder_weights(), mel[1:2])
    wav = engine.forward(mel.cuda()).cpu()
    err = tensor_rel_err(wav[1:2], ref)
    print(f"vocoder B=2 T=512, utterance 1: {err:.2e}")
    assert err < 1e-3
    again = engine.forward(mel.cuda()).cpu()                 # graph replay: bit-identical
    assert torch.equal(again, wav)


def test_dropin_generator_loads_weight_normed_checkpoint():
    """hifigan.Generator as get_vocoder uses it (DEX-TTS/src/utils.py:251-281): load a weight-normed state dict, eval, remove_weight_norm,
    to(device), call.  The checkpoint is synthesised: weight_v = the seeded weight, weight_g = 2 * its norm, so the folded weight is
    2 * the seeded weight for every convolution (bias unchanged)."""
    from dexb200.hifigan import AttrDict, Generator
    w = V.synth_vocoder_weights()
    sd = {}
    for name, t in w.items():
        if name.endswith(".bias"):
            sd[name] = t
        else:
            sd[name + "_v"] = t
            sd[name + "_g"] = 2.0 * torch.norm_except_dim(t, 2, 0)
    voc = Generator(AttrDict(CFG))
    assert set(voc.state_dict().keys()) == set(sd.keys())
    voc.load_state_dict(sd)
    voc.eval()
    voc.remove_weight_norm()
    voc.to("cuda")
    gen = torch.Generator()
    gen.manual_seed(5)
    mel = torch.randn(1, 80, 21, generator=gen) * 1.5 - 4.0
    w2 = {k: (v if k.endswith(".bias") else 2.0 * v) for k, v in w.items()}
    with torch.no_grad():
        ref = V.hifigan_generator(w2, mel)
    wav = voc(mel.cuda()).cpu()
    err = tensor_rel_err(wav, ref)
    print(f"drop-in Generator: {err:.2e}")
    assert err < 1e-3
