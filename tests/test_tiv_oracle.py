"""Pin the oracle restatement of the TIV encoder (oracle/dex_oracle.py: tiv_encoder) against outputs of the unmodified reference
TIVEncoder (tests/golden/tiv_*.npz, made by oracle/make_golden_tiv.py in the build container), and the drop-in module's
state-dict layout against the reference's key list stored in the same fixtures."""
import glob
import os

import numpy as np
import pytest
import torch

import dex_oracle as O
from dexb200.synth import synth_ref_mel, synth_tiv_weights, tiv_manifest
from parity import tensor_rel_err

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "tiv_*.npz")))


def load_case(path):
    g = np.load(path)
    B, T, ragged, seed = [int(v) for v in g["meta"]]
    return g, synth_ref_mel(B, T, seed=seed, ragged=bool(ragged))


def test_golden_present():
    assert len(GOLD) >= 2


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_reference(path):
    g, inp = load_case(path)
    with torch.no_grad():
        out, skips = O.tiv_encoder(synth_tiv_weights(), inp["ref"], inp["mask"])
    assert len(skips) == 6
    assert tensor_rel_err(out, torch.from_numpy(g["out"])) < 2e-6        # same ATen kernels, same op order
    for i, s in enumerate(skips):
        assert tensor_rel_err(s, torch.from_numpy(g[f"skip{i}"])) < 2e-6, i
    # padded frames of every skip are exactly zero (x * mask, ref_encoder.py:101)
    pad = 1.0 - inp["mask"]
    assert float((skips[-1] * pad).abs().max()) == 0.0


def test_module_state_dict_is_the_reference_layout():
    from dexb200.model import TIVEncoder
    m = TIVEncoder(c_in=80, c_out=64, num_layer=6, c_h=128)
    keys = [str(k) for k in np.load(GOLD[0])["keys"]]                     # reference TIVEncoder.state_dict().keys(), in order
    assert list(m.state_dict().keys()) == keys
    assert keys == [n for n, _, _ in tiv_manifest()]
    m.load_state_dict(synth_tiv_weights(prefix=""), strict=True)
    sd = m.state_dict()
    assert sd["in_conv.conv.weight"].shape == (128, 80, 3) and sd["out_conv.bn.running_var"].shape == (64,)
    assert sd["in_conv.bn.num_batches_tracked"].dtype == torch.long
    # default construction = BatchNorm1d defaults
    m2 = TIVEncoder(c_in=80, c_out=64, num_layer=6, c_h=128)
    assert float(m2.state_dict()["in_conv.bn.running_var"].min()) == 1.0 and float(m2.state_dict()["in_conv.bn.bias"].abs().max()) == 0.0


def test_module_refuses_cpu_and_training():
    from dexb200.model import TIVEncoder
    m = TIVEncoder(c_in=80, c_out=64, num_layer=6, c_h=128)
    with pytest.raises(NotImplementedError):
        m.train()(torch.zeros(1, 80, 8), torch.ones(1, 1, 8))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            m.eval()(torch.zeros(1, 80, 8), torch.ones(1, 1, 8))
