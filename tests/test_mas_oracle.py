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
