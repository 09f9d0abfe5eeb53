"""Pin the oracle restatement of the TV encoder (oracle/dex_oracle.py: tv_encoder) against outputs of the unmodified reference
TVEncoder (tests/golden/tv_*.npz, made by oracle/make_golden_tv.py in the build container), and the drop-in module's
state-dict layout against the reference's key list stored in the same fixtures."""
import glob
import os

import numpy as np
import pytest
import torch

import dex_oracle as O
from dexb200.synth import synth_ref_mel, synth_tv_weights, tv_manifest
from parity import tensor_rel_err

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "tv_*.npz")))


def test_golden_present():
    assert len(GOLD) >= 2


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_reference(path):
    g = np.load(path)
    B, T, ragged, seed = [int(v) for v in g["meta"]]
    inp = synth_ref_mel(B, T, seed=seed, ragged=bool(ragged))
    with torch.no_grad():
        z_before, z_dec, loss, idx = O.tv_encoder(synth_tv_weights(), inp["ref"], inp["mask"], return_indices=True)
    assert np.array_equal(idx.numpy(), g["idx"])                              # the codes the reference's argmin picked
    assert tensor_rel_err(z_before, torch.from_numpy(g["z_before"])) < 2e-6  # same ATen kernels, same op order
    assert tensor_rel_err(z_dec, torch.from_numpy(g["z_dec"])) < 2e-6
    assert abs(float(loss) - float(g["vq_loss"])) < 1e-6 * float(g["vq_loss"])
    assert len(set(g["idx"].ravel().tolist())) >= 10                          # the fixture exercises many codes


def test_module_state_dict_is_the_reference_layout():
    from dexb200.model import TVEncoder
    m = TVEncoder(c_in=80, c_out=192, c_out_g=192, num_layer=6, c_h=128, n_emb=512, commit_w=0.25)
    keys = [str(k) for k in np.load(GOLD[0])["keys"]]                         # reference TVEncoder.state_dict().keys(), in order
    assert list(m.state_dict().keys()) == keys
    assert keys == [n for n, _, _ in tv_manifest()]
    m.load_state_dict(synth_tv_weights(prefix=""), strict=True)
    sd = m.state_dict()
    assert sd["vq.embedding"].shape == (512, 192) and sd["proj_0.proj.weight"].shape == (192, 192, 1)
    with pytest.raises(NotImplementedError):
        m.train()(torch.zeros(1, 80, 8), torch.ones(1, 1, 8))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            m.eval()(torch.zeros(1, 80, 8), torch.ones(1, 1, 8))
