"""Pin the oracle restatements of the LF0 encoder and of the style fusion of DeXTTS.forward (oracle/dex_oracle.py: lf0_encoder,
style_fusion) against outputs of the unmodified reference modules (tests/golden/lf0_*.npz, made by oracle/make_golden_lf0.py in
the build container), and the drop-in module's state-dict layout against the reference's key list."""
import glob
import os

import numpy as np
import pytest
import torch

import dex_oracle as O
from dexb200.synth import (lf0_manifest, synth_conv_sty_weights, synth_lf0, synth_lf0_weights, synth_ref_mel, synth_tv_weights)
from parity import tensor_rel_err

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "lf0_*.npz")))


def oracle_case(path):
    g = np.load(path)
    B, T, ragged, seed = [int(v) for v in g["meta"]]
    inp = synth_lf0(B, T, seed=seed, ragged=bool(ragged))
    sty = synth_ref_mel(B, T, seed=seed + 100, ragged=bool(ragged))
    w = dict(synth_lf0_weights())
    w.update(synth_tv_weights())
    w.update(synth_conv_sty_weights())
    with torch.no_grad():
        le, ld = O.lf0_encoder(w, inp["lf0"], inp["mask"])
        zb, zd, _ = O.tv_encoder(w, sty["ref"], sty["mask"])
        se, sd = O.style_fusion(w, zb, zd, le, ld, sty["mask"], inp["mask"])
    return g, inp, sty, dict(lf0_enc=le, lf0_dec=ld, sty_enc=se, sty_dec=sd, z_before=zb, z_dec=zd)


def test_golden_present():
    assert len(GOLD) >= 2


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_reference(path):
    g, _, _, o = oracle_case(path)
    # the GRU is restated step by step (nn.GRU's fused CPU kernel orders its sums differently): fp32 reassociation noise only
    for k in ("lf0_enc", "lf0_dec", "sty_enc", "sty_dec"):
        ref = torch.from_numpy(g[k])
        assert o[k].shape == ref.shape, k
        assert tensor_rel_err(o[k], ref) < 2e-5, k


def test_module_state_dict_is_the_reference_layout():
    from dexb200.model import LF0Encoder
    m = LF0Encoder(c_h=192, c_out=192, c_out_g=192, num_layer=2, c_in=1)
    keys = [str(k) for k in np.load(GOLD[0])["keys"]]                         # reference LF0Encoder.state_dict().keys(), in order
    assert list(m.state_dict().keys()) == keys
    assert keys == [n for n, _, _ in lf0_manifest()]
    m.load_state_dict(synth_lf0_weights(prefix=""), strict=True)
    assert m.state_dict()["rnn_layer.weight_hh_l1_reverse"].shape == (288, 96)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            m.eval()(torch.zeros(1, 8), torch.ones(1, 1, 8))
