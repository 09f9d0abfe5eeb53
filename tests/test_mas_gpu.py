"""GPU parity of Monotonic Alignment Search (dexb_mas_maximum_path through the C ABI, drop-in ``model.monotonic_align.maximum_path``)
against the output of the reference's own Cython kernel (tests/golden/mas_*.npz) and the CPU oracle: bit-exact, the result is a path
of zeros and ones.  Written after this round's GPU budget was spent: NOT YET RUN on a B200 (run_last); the kernel's arithmetic is
replayed statement by statement against the fixtures in tests/test_mas_oracle.py."""
import glob
import os

import numpy as np
import pytest
import torch

import mas_oracle as MO
from test_mas_oracle import check_monotonic, load_case
from make_golden_mas import synth_mas

pytestmark = [pytest.mark.gpu, pytest.mark.run_last(2)]      # after the text-side items: this kernel has never run on a GPU

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "mas_*.npz")))


def run_gpu(value, mask):
    from dexb200.model.monotonic_align import maximum_path
    out = maximum_path(torch.from_numpy(value).cuda(), torch.from_numpy(mask).cuda())
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_mas_matches_reference_kernel_bit_exactly(path):
    value, mask, ref = load_case(path)
    out = run_gpu(value, mask)
    assert out.shape == ref.shape and np.array_equal(out, ref)


@pytest.mark.parametrize("B,Tx,Ty", [(8, 128, 600), (2, 512, 2400), (1, 1, 1), (3, 33, 33)])
def test_mas_matches_oracle_at_baseline_sizes(B, Tx, Ty):
    """Text / mel lengths of BASELINE.json's configs (128 x ~600: C2; 512 x ~2400: C5), one cell, and the square case Tx = Ty."""
    value, mask = synth_mas(B, Tx, Ty, seed=7 + Tx)
    out = run_gpu(value, mask)
    assert np.array_equal(out, MO.maximum_path(value, mask))
    check_monotonic(out, mask)
