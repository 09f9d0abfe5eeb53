"""GPU parity of the duration / alignment glue (dexb_align_lengths / dexb_align_expand through the C ABI, drop-in
``model.align_durations``) against the reference fixtures (tests/golden/align_*.npz) and the CPU oracle: bit-exact, this is
index / byte work (run on a B200: profiles/r01_pytest_gpu_v31_align.log)."""
import glob
import os

import numpy as np
import pytest
import torch

import dex_oracle as O
from dexb200.synth import synth_align_inputs
from test_align_oracle import check_alignment_properties, load_case

pytestmark = pytest.mark.gpu

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "align_*.npz")))


def run_gpu(inp, length_scale=1.0, want_attn=True):
    from dexb200.model import align_durations
    out = align_durations(inp["logw"].cuda(), inp["x_mask"].cuda(), inp["mu_x"].cuda(), length_scale, want_attn=want_attn)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_align_matches_reference_fixture_bit_exactly(path):
    g, inp, length_scale, attn_ref, y_max_ref, Ty_ref = load_case(path)
    mu_y, y_mask, attn, y_lengths, y_max = run_gpu(inp, length_scale)
    assert y_max == y_max_ref and mu_y.shape[-1] == Ty_ref and attn.shape == attn_ref.shape
    assert np.array_equal(y_lengths.cpu().numpy(), g["y_lengths"])
    assert np.array_equal(y_mask.cpu().numpy(), g["y_mask"])
    assert np.array_equal(attn.cpu().numpy(), attn_ref)
    assert np.array_equal(mu_y.cpu().numpy(), g["mu_y"])


@pytest.mark.parametrize("B,Tx,ragged", [(8, 128, True), (8, 512, True), (32, 128, False), (1, 1, False)])
def test_align_matches_oracle_at_baseline_sizes(B, Tx, ragged):
    """Text lengths of BASELINE.json's configs (128: C2 / C3, 512: C5) and the smallest input."""
    inp = synth_align_inputs(B, Tx, seed=70 + B + Tx, ragged=ragged, mean_dur=4.0)
    mu_y, y_mask, attn, y_lengths, y_max = run_gpu(inp)
    r_mu, r_mask, r_attn, r_len, r_max = O.align_durations(inp["logw"], inp["x_mask"], inp["mu_x"])
    assert y_max == r_max and torch.equal(y_lengths.cpu(), r_len)
    assert torch.equal(y_mask.cpu(), r_mask) and torch.equal(attn.cpu(), r_attn) and torch.equal(mu_y.cpu(), r_mu)
    check_alignment_properties(attn.cpu(), y_mask.cpu(), y_lengths.cpu(), inp["x_mask"])


def test_align_without_attn_and_zero_durations():
    inp = synth_align_inputs(2, 40, seed=5, ragged=True)
    full = run_gpu(inp)
    lean = run_gpu(inp, want_attn=False)
    assert lean[2] is None and torch.equal(full[0], lean[0]) and torch.equal(full[1], lean[1])
    # exp(logw) underflows to 0 everywhere: clamp_min(., 1) (tts.py:57) -> one frame nobody owns
    z = dict(logw=torch.full((1, 1, 5), -200.0), x_mask=torch.ones(1, 1, 5), mu_x=torch.randn(1, 80, 5))
    mu_y, y_mask, attn, y_lengths, y_max = run_gpu(z)
    assert y_lengths.tolist() == [1] and y_max == 1 and mu_y.shape[-1] == 4
    assert float(attn.abs().max()) == 0.0 and float(mu_y.abs().max()) == 0.0 and y_mask.flatten().tolist() == [1, 0, 0, 0]


def test_align_feeds_the_decoder_shapes():
    """mu_y / y_mask have the decoder's layout: (B, 80, Ty_) and (B, 1, Ty_) with Ty_ a multiple of 4 (fix_len_compatibility)."""
    inp = synth_align_inputs(2, 30, seed=11, ragged=True, mean_dur=3.0)
    mu_y, y_mask, attn, y_lengths, y_max = run_gpu(inp)
    assert mu_y.shape[-1] % 4 == 0 and mu_y.shape == (2, 80, y_mask.shape[-1]) and mu_y.is_contiguous()
    assert attn.shape == (2, 1, 30, mu_y.shape[-1]) and int(y_lengths.max()) == y_max
    assert float((mu_y * (1 - y_mask)).abs().max()) == 0.0
