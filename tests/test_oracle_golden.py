"""Pin the oracle restatement against outputs of the unmodified reference (tests/golden/*.npz, made by
oracle/make_golden.py in the build container)."""
import glob
import os

import numpy as np
import pytest
import torch

import dex_oracle as O
from dexb200.manifest import DecoderCfg
from dexb200.synth import synth_decoder_weights, synth_inputs
from parity import per_bin_violation, tensor_rel_err

GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
              if os.path.basename(p).startswith(("dex_", "gedex_", "libri_dex_")))       # libri_*: LibriTTS decoder sizes (dim 128 / hidden 384)


def load_case(path):
    g = np.load(path)
    meta = [int(v) for v in g["meta"]]
    B, T, Ts, steps, ragged, live, seed = meta[:7]
    variant = str(g["variant"])
    dims = dict(dim=int(g["dims"][0]), hidden=int(g["dims"][1])) if "dims" in g.files else {}
    cfg = DecoderCfg.make(variant, n_spks=meta[7] if len(meta) > 7 else None, **dims)     # 8th entry: multi-speaker GeDEX-TTS
    w = synth_decoder_weights(cfg, seed=100, live=bool(live))
    inp = synth_inputs(cfg, B, T, Ts=max(Ts, 1), seed=seed, ragged=bool(ragged))
    cond = None
    if variant == "dex":
        cond = dict(sty=inp["sty"], sty_lengths=inp["sty_lengths"], ref_skips=inp["ref_skips"])
    elif cfg.n_spks > 1:
        cond = dict(spk=inp["spk"])
    return g, cfg, w, inp, cond, steps


def test_golden_present():
    assert len(GOLD) >= 6


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_reference(path):
    g, cfg, w, inp, cond, steps = load_case(path)
    ocfg = O.make_cfg(cfg.variant, n_spks=cfg.n_spks, dim=cfg.dim, hidden=cfg.hidden)
    with torch.no_grad():
        y = O.reverse_diffusion(w, ocfg, inp["z"], inp["mask"], inp["mu"], steps,
                                temperature=float(g["temperature"]), cond=cond)
    y_ref = torch.from_numpy(g["y"])
    assert y.shape == y_ref.shape
    # same ATen kernels, same op order -> essentially bit-equal; allow fp32 reassociation noise only
    assert per_bin_violation(y, y_ref) < 2e-5


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_intermediates(path):
    g, cfg, w, inp, cond, steps = load_case(path)
    ocfg = O.make_cfg(cfg.variant, n_spks=cfg.n_spks, dim=cfg.dim, hidden=cfg.hidden)
    taps = {}
    ts = O.sigma_schedule(steps)
    x0 = (inp["z"] / float(g["temperature"]) + inp["mu"]) * ts[0]
    with torch.no_grad():
        O.edm_precond(w, ocfg, x0, ts[0], inp["mask"], inp["mu"], cond=cond, taps=taps)
    checked = 0
    wst = int(g["tap_wstride"]) if "tap_wstride" in g.files else 1      # T >= 256 fixtures are also strided along time
    for k in ("tv_out", "tiv_out", "dit_out", "up_out"):
        if "tap_" + k in g.files:
            ref = torch.from_numpy(g["tap_" + k])
            assert tensor_rel_err(taps[k][:, ::16, :, ::wst], ref) < 2e-5, k
            checked += 1
    assert checked >= 2
