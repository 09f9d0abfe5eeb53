"""Pin the oracle restatement of Monotonic Alignment Search (oracle/mas_oracle.py; DEX-TTS/model/monotonic_align/core.pyx:9-37,
__init__.py:8-25) against the output of the reference's own Cython kernel compiled in the build container (tests/golden/mas_*.npz,
oracle/make_golden_mas.py): bit-exact, it is a path of zeros and ones.  SURVEY.md §8f rank 4 (training side)."""
import glob
import os
import sys

import numpy as np
import pytest

import mas_oracle as MO

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from make_golden_mas import synth_mas  # noqa: E402

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "mas_*.npz")))


def load_case(path):
    g = np.load(path)
    B, Tx, Ty, seed = [int(v) for v in g["meta"]]
    shape = tuple(int(v) for v in g["shape"])
    ref = np.unpackbits(g["path"], axis=-1, count=shape[-1]).reshape(shape).astype(np.float32)
    value, mask = synth_mas(B, Tx, Ty, seed)
    return value, mask, ref


def check_monotonic(path, mask):
    """One token per valid frame, token index non-decreasing by at most one per frame, first frame on token 0, last on the last."""
    t_x = mask.sum(1)[:, 0].astype(int)
    t_y = mask.sum(2)[:, 0].astype(int)
    for b in range(path.shape[0]):
        p = path[b, :, :t_y[b]]
        assert np.array_equal(p.sum(0), np.ones(t_y[b])) and path[b, :, t_y[b]:].sum() == 0 and path[b, t_x[b]:].sum() == 0
        owner = p.argmax(0)
        assert owner[0] == 0 and owner[-1] == t_x[b] - 1 and set(np.diff(owner)) <= {0, 1}


def test_golden_present():
    assert len(GOLD) >= 3


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_reference_kernel_bit_exactly(path):
    value, mask, ref = load_case(path)
    out = MO.maximum_path(value, mask)
    assert out.shape == ref.shape and np.array_equal(out, ref)
    check_monotonic(out, mask)


def test_oracle_path_is_optimal_on_a_small_case():
    """Exhaustive check: among all monotonic paths of a 4 x 7 score matrix the returned one has the maximum total score."""
    import itertools
    rng = np.random.default_rng(5)
    value = rng.standard_normal((1, 4, 7)).astype(np.float32)
    mask = np.ones_like(value)
    out = MO.maximum_path(value, mask)[0]
    best = max(sum(value[0, x, y] for y, x in enumerate(np.cumsum((0,) + steps)))
               for steps in itertools.product((0, 1), repeat=6) if sum(steps) == 3)
    assert abs(float((out * value[0]).sum()) - float(best)) < 1e-5


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_kernel_arithmetic_reproduces_reference(path):
    """csrc/mas.cu replayed in numpy float32, statement by statement: only the previous column is kept (two buffers), the backtrack
    decision is taken in the forward pass and stored per (y, x), the walk back reads the stored decisions.  Checks the algorithm the
    kernel implements against the reference kernel's output on a box without a GPU; the kernel itself: tests/test_mas_gpu.py."""
    f32 = np.float32
    value, mask, ref = load_case(path)
    B, Tx, Ty = value.shape
    out = np.zeros_like(ref)
    for b in range(B):
        t_x = int((mask[b, :, 0] != 0).sum())
        t_y = int((mask[b, 0, :] != 0).sum())
        dec = np.zeros((Ty, Tx), np.uint8)
        prev, cur = np.full(Tx, np.nan, f32), np.full(Tx, np.nan, f32)          # NaN: a read outside the previous band would show
        for y in range(t_y):
            for x in range(max(0, t_x + y - t_y), min(t_x, y + 1)):
                raw = f32(value[b, x, y] * mask[b, x, y])
                v_cur = f32(-1e9) if x == y else prev[x]
                v_prev = (f32(0) if y == 0 else f32(-1e9)) if x == 0 else prev[x - 1]
                assert not (np.isnan(v_cur) or np.isnan(v_prev))
                cur[x] = f32(max(v_cur, v_prev) + raw)
                dec[y, x] = 1 if (x == y or v_cur < v_prev) else 0
            prev, cur = cur, prev
            cur[:] = np.nan
        index = t_x - 1
        for y in range(t_y - 1, -1, -1):
            out[b, index, y] = 1
            if index != 0 and y > 0 and dec[y, index]:
                index -= 1
    assert np.array_equal(out, ref)
