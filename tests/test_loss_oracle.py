"""Pin the oracle restatement of the training loss (oracle/dex_oracle.py: edm_loss; EDMLoss.forward, DEX-TTS/model/edm.py:31-68, through
Diffusion.forward(infer=False), diffusion.py:252-254) against the value the unmodified reference returns on the same seeded weights,
inputs and Gaussian draws (tests/golden/loss_*.npz, oracle/make_golden_loss.py).  SURVEY.md §8f rank 4; forward value only, no CUDA side."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

import dex_oracle as O
from dexb200.manifest import DecoderCfg
from dexb200.synth import synth_decoder_weights, synth_inputs

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from make_golden_loss import loss_draws  # noqa: E402

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "loss_*.npz")))


def test_golden_present():
    assert len(GOLD) >= 2


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_loss_oracle_matches_reference(path):
    g = np.load(path)
    variant = str(g["variant"])
    B, T, Ts, seed = [int(v) for v in g["meta"]]
    cfg = DecoderCfg.make(variant)
    w = synth_decoder_weights(cfg, seed=100, live=True)
    inp = synth_inputs(cfg, B, T, Ts=max(Ts, 1), seed=seed, ragged=True)
    rnd, noise = loss_draws(B, T, seed + 1)
    cond = dict(sty=inp["sty"], sty_lengths=inp["sty_lengths"], ref_skips=inp["ref_skips"]) if variant == "dex" else None
    with torch.no_grad():
        loss = O.edm_loss(w, O.make_cfg(variant), inp["z"], inp["mask"], inp["mu"], rnd, noise, cond=cond)
    ref = float(g["loss"])
    print(f"{os.path.basename(path)}: loss {float(loss):.6f} vs reference {ref:.6f}")
    assert abs(float(loss) - ref) <= 2e-5 * abs(ref)
