"""Pin the oracle restatement of the text encoder (oracle/text_oracle.py: prenet, RetNet-as-softmax-attention with rotary q / k,
swish gate and AdaLN(style), duration predictor; DEX-TTS/model/text_encoder.py:129-142) against outputs of the unmodified reference
TextEncoder (tests/golden/text_*.npz, made by oracle/make_golden_text.py in the build container).  SURVEY.md §8f rank 2: the oracle
of the next stage to be built -- there is no CUDA side for it yet, so there is no GPU test beside this file."""
import glob
import os

import numpy as np
import pytest
import torch

import dex_oracle as O
import text_oracle as TO
from dexb200.synth import synth_text, synth_text_weights, text_manifest
from parity import tensor_rel_err

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "text_*.npz")))
TOL = 2e-5      # measured 0 (bit-exact) in the build container; the bound covers a BLAS that groups sums differently with other thread counts


def oracle_case(path):
    g = np.load(path)
    B, Tx, ragged, seed, dex, n_spks = [int(v) for v in g["meta"]]
    inp = synth_text(B, Tx, seed=seed, ragged=bool(ragged))
    taps = {}
    spk = torch.randn(B, 64, generator=torch.Generator().manual_seed(seed + 7)) if n_spks > 1 else None
    with torch.no_grad():                                   # dex = 0: GeDEX-TTS's encoder (no AdaLN, no style argument)
        mu, logw, x_mask = TO.text_encoder(synth_text_weights(adaln=bool(dex), spk_emb_dim=64 if n_spks > 1 else 0), inp["x"],
                                           inp["x_lengths"], inp["sty"] if dex else None, taps=taps, spk=spk)
    return g, inp, mu, logw, x_mask, taps


def test_golden_present():
    assert len(GOLD) >= 4


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_manifest_is_the_reference_state_dict_layout(path):
    g = np.load(path)
    assert [n for n, _, _ in text_manifest(adaln=bool(g["meta"][4]), spk_emb_dim=64 if g["meta"][5] > 1 else 0)] == [str(k) for k in g["keys"]]


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_reference(path):
    g, inp, mu, logw, x_mask, taps = oracle_case(path)
    assert np.array_equal(x_mask.numpy(), g["x_mask"])
    errs = {k: tensor_rel_err(v, torch.from_numpy(g[k])) for k, v in
            (("prenet", taps["prenet"]), ("layer0", taps["layer0"]), ("layer7", taps["layer7"]), ("mu", mu), ("logw", logw))}
    print(f"text fixture {os.path.basename(path)}: " + " ".join(f"{k} {e:.2e}" for k, e in errs.items()))
    assert max(errs.values()) < TOL, errs
    # padded tokens: mu and logw are exactly zero (text_encoder.py:138, 95)
    pad = 1 - x_mask
    assert float((mu * pad).abs().max()) == 0.0 and float((logw * pad).abs().max()) == 0.0


def test_style_vector_reaches_the_output():
    """The AdaLN weights are zero-initialised upstream (base.py:174-178); the synthetic weights re-draw them, so sty matters."""
    inp = synth_text(1, 12, seed=3)
    w = synth_text_weights()
    with torch.no_grad():
        a = TO.text_encoder(w, inp["x"], inp["x_lengths"], inp["sty"])[0]
        b = TO.text_encoder(w, inp["x"], inp["x_lengths"], inp["sty"] + 1.0)[0]
    assert tensor_rel_err(a, b) > 1e-2


def test_text_oracle_feeds_the_alignment_glue():
    """TextEncoder -> duration / alignment glue (tts.py:52-68): shapes and masks line up; every valid token with duration >= 1
    owns at least one frame."""
    inp = synth_text(2, 21, seed=4, ragged=True)
    with torch.no_grad():
        mu_x, logw, x_mask = TO.text_encoder(synth_text_weights(), inp["x"], inp["x_lengths"], inp["sty"])
        mu_y, y_mask, attn, y_lengths, y_max = O.align_durations(logw, x_mask, mu_x)
    assert mu_y.shape == (2, 80, y_mask.shape[-1]) and attn.shape == (2, 1, 21, y_mask.shape[-1])
    owned = attn.squeeze(1).sum(-1)
    assert torch.equal(owned, torch.ceil(torch.exp(logw) * x_mask).squeeze(1))
