"""The LibriTTS sizes of DEX-TTS (DEX-TTS/config/LibriTTS/base.yaml: decoder dim 128, DiT hidden 384 / head dim 192, 256-wide text, TV,
LF0 and TIV encoders) -- the third shipped config of the reference beside VCTK (DEX-TTS) and LJSpeech (GeDEX-TTS).

CPU side: the oracle chain (oracle/tts_oracle.py) against outputs of the UNMODIFIED reference ``DeXTTS`` built from that yaml
(tests/golden/libritts_dex_*.npz, oracle/make_golden_tts.py), and the drop-in's ``state_dict`` against the reference's key list.
The CUDA side is tests/test_libritts_gpu.py."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

import dex_oracle as O
import tts_oracle as TT
from dexb200.synth import LIBRITTS_MODEL_CFG, reference_state_dict, seeded_noise, synth_tts_weights
from parity import REL_TOL, per_bin_violation

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from make_golden_tts import synth_tts_inputs  # noqa: E402

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "libritts_dex_*.npz")))
DCFG = dict(dim=128, hidden=384)


def build_libritts():
    from dexb200.model import DeXTTS
    model = DeXTTS(LIBRITTS_MODEL_CFG)
    w = synth_tts_weights("dex", dataset="LibriTTS")
    model.load_state_dict(reference_state_dict(w), strict=True)
    return model.eval(), w


def test_golden_present():
    assert len(GOLD) >= 2


def test_state_dict_is_the_reference_models():
    g = np.load(GOLD[0])
    assert str(g["dataset"]) == "LibriTTS"
    model, _ = build_libritts()
    sd = model.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    assert [list(v.shape) for v in sd.values()] == [[int(n) for n in s.split(",") if n] for s in g["shapes"]]


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_forward_oracle_matches_reference_forward(path):
    g = np.load(path)
    B, Tx, Ts, steps, ragged, seed = [int(v) for v in g["meta"]]
    temperature, length_scale = [float(v) for v in g["scale"]]
    inp = synth_tts_inputs("dex", B, Tx, Ts, seed, bool(ragged))
    w = synth_tts_weights("dex", dataset="LibriTTS")
    with torch.no_grad():
        enc_out, dec_out, attn = TT.dextts_forward(w, inp["x"], inp["x_lengths"], inp["ref"], inp["ref_lengths"], inp["ref"],
                                                   inp["ref_lengths"], inp["lf0"], inp["lf0_lengths"], steps, seeded_noise(seed + 3),
                                                   temperature, length_scale, decoder_cfg=O.make_cfg("dex", **DCFG))
    shape = tuple(int(v) for v in g["attn_shape"])
    attn_ref = np.unpackbits(g["attn"], axis=-1, count=shape[-1]).astype(np.float32).reshape(shape)
    assert attn.shape == shape and np.array_equal(attn.numpy(), attn_ref)
    v_enc = per_bin_violation(enc_out, torch.from_numpy(g["enc_out"]))
    v_dec = per_bin_violation(dec_out, torch.from_numpy(g["dec_out"]))
    print(f"{os.path.basename(path)}: enc_out {v_enc:.2e} dec_out {v_dec:.2e}")
    assert v_enc < 2e-5 and v_dec < REL_TOL
