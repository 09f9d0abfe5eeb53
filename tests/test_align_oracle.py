"""Pin the oracle restatement of the duration / alignment glue (oracle/dex_oracle.py: align_durations; DEX-TTS/model/tts.py:55-68)
against outputs of the reference's own model.utils functions (tests/golden/align_*.npz, made by oracle/make_golden_align.py in the
build container), bit-exactly, plus the size-independent properties of a hard monotonic alignment and the host-side helpers of
the drop-in ``model.utils``."""
import glob
import os

import numpy as np
import pytest
import torch

import dex_oracle as O
from dexb200.synth import synth_align_inputs

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "align_*.npz")))


def load_case(path):
    g = np.load(path)
    B, Tx, ragged, seed, y_max, Ty = [int(v) for v in g["meta"]]
    length_scale, mean_dur = [float(v) for v in g["scale"]]
    inp = synth_align_inputs(B, Tx, seed=seed, ragged=bool(ragged), mean_dur=mean_dur)
    shape = tuple(int(v) for v in g["attn_shape"])
    attn = np.unpackbits(g["attn"], axis=-1, count=shape[-1]).astype(np.float32).reshape(shape)
    return g, inp, length_scale, attn, y_max, Ty


def check_alignment_properties(attn, y_mask, y_lengths, x_mask):
    """Every valid output frame is owned by exactly one unmasked token, owners are non-decreasing in time, padding owns nothing."""
    a = attn.squeeze(1)
    B, Tx, Ty = a.shape
    col = a.sum(1)
    assert torch.equal(col, y_mask.squeeze(1))
    assert torch.equal(y_mask.squeeze(1).sum(1).long(), y_lengths.clamp(max=Ty))
    assert float((a * (1 - x_mask.squeeze(1))[:, :, None]).abs().max()) == 0.0
    owner = a.argmax(1)
    for b in range(B):
        n = int(y_lengths[b])
        assert bool((owner[b, 1:n] >= owner[b, :n - 1]).all())


def test_golden_present():
    assert len(GOLD) >= 3


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_reference_bit_exactly(path):
    g, inp, length_scale, attn_ref, y_max_ref, Ty_ref = load_case(path)
    mu_y, y_mask, attn, y_lengths, y_max = O.align_durations(inp["logw"], inp["x_mask"], inp["mu_x"], length_scale)
    assert y_max == y_max_ref and mu_y.shape[-1] == Ty_ref
    assert np.array_equal(y_lengths.numpy(), g["y_lengths"])
    assert np.array_equal(y_mask.numpy(), g["y_mask"])
    assert np.array_equal(attn.numpy(), attn_ref)
    assert np.array_equal(mu_y.numpy(), g["mu_y"])
    if length_scale == 1.0:
        check_alignment_properties(attn, y_mask, y_lengths, inp["x_mask"])


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_kernel_index_arithmetic_reproduces_reference(path):
    """The arithmetic of csrc/align.cu replayed in numpy float32, statement by statement (k_align_len: sequential un-fused
    mul / add, truncation; k_align_expand: binary search for the first cum[i] > t, one-hot store, gather) -- checks the algorithm
    the kernels implement against the reference fixtures on a box without a GPU.  The kernels themselves: tests/test_align_gpu.py."""
    f32 = np.float32
    g, inp, length_scale, attn_ref, y_max, Ty = load_case(path)
    logw, xm, mu_x = inp["logw"].numpy()[:, 0], inp["x_mask"].numpy()[:, 0], inp["mu_x"].numpy()
    B, Tx = logw.shape
    cum, y_len = np.zeros((B, Tx), f32), np.zeros(B, np.int64)
    for b in range(B):
        c = f32(0)
        for i in range(Tx):
            w = f32(np.exp(logw[b, i])) * xm[b, i]
            c = f32(c + f32(np.ceil(w) * f32(length_scale)))
            cum[b, i] = c
        y_len[b] = int(max(c, f32(1)))
    assert np.array_equal(y_len, g["y_lengths"])
    attn, mu_y, y_mask = np.zeros((B, Tx, Ty), f32), np.zeros((B, 80, Ty), f32), np.zeros((B, Ty), f32)
    for b in range(B):
        for t in range(Ty):
            ym = f32(1) if t < y_len[b] else f32(0)
            y_mask[b, t] = ym
            lo = int(np.searchsorted(cum[b], f32(t), side="right"))         # first i with t < cum[i]
            if lo < Tx:
                a = xm[b, lo] * ym
                attn[b, lo, t] = a
                mu_y[b, :, t] = a * mu_x[b, :, lo]
    assert np.array_equal(attn, attn_ref[:, 0]) and np.array_equal(mu_y, g["mu_y"]) and np.array_equal(y_mask, g["y_mask"][:, 0])


def test_oracle_gather_equals_matmul_at_full_size():
    """BASELINE.json's long-form text length (C5: 512 phonemes): `attn^T mu_x` is a gather of mu_x columns."""
    inp = synth_align_inputs(4, 512, seed=9, ragged=True, mean_dur=4.0)
    mu_y, y_mask, attn, y_lengths, y_max = O.align_durations(inp["logw"], inp["x_mask"], inp["mu_x"])
    check_alignment_properties(attn, y_mask, y_lengths, inp["x_mask"])
    owner = attn.squeeze(1).argmax(1)
    gathered = torch.gather(inp["mu_x"], 2, owner[:, None, :].expand(-1, 80, -1)) * y_mask
    assert torch.equal(gathered, mu_y)


def test_oracle_all_durations_zero_gives_length_one():
    """exp(logw) underflows to 0 -> sum 0 -> clamp_min(., 1) (tts.py:57): one output frame that no token owns."""
    logw = torch.full((1, 1, 5), -200.0)
    x_mask = torch.ones(1, 1, 5)
    mu_y, y_mask, attn, y_lengths, y_max = O.align_durations(logw, x_mask, torch.randn(1, 80, 5))
    assert y_lengths.tolist() == [1] and y_max == 1 and mu_y.shape[-1] == 4
    assert float(attn.abs().max()) == 0.0 and float(mu_y.abs().max()) == 0.0 and y_mask.flatten().tolist() == [1, 0, 0, 0]


def test_host_helpers_match_reference_semantics():
    from dexb200.model.utils import fix_len_compatibility, sequence_mask
    for n in range(1, 40):
        ref = n
        while ref % 4:
            ref += 1
        assert fix_len_compatibility(n) == ref
    assert fix_len_compatibility(9, 3) == 16 and fix_len_compatibility(8, 3) == 8
    m = sequence_mask(torch.tensor([3, 0, 5]), 6)
    assert m.dtype == torch.bool and m.long().sum(1).tolist() == [3, 0, 5] and bool(m[0, :3].all())
    assert sequence_mask(torch.tensor([2, 4])).shape == (2, 4)


def test_align_durations_refuses_cpu_tensors():
    from dexb200.model import align_durations
    inp = synth_align_inputs(1, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        align_durations(inp["logw"], inp["x_mask"], inp["mu_x"])
