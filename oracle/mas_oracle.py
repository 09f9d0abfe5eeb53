"""CPU oracle for Monotonic Alignment Search, the reference's only native (Cython) component -- training side, SURVEY.md §8f rank 4.

TEST INFRASTRUCTURE ONLY -- never imported by the product path (``dex-tts_b200/``); see the header of ``dex_oracle.py``.

numpy float32 restatement of

    maximum_path          DEX-TTS/model/monotonic_align/__init__.py:8-25   (value * mask, lengths from the mask's first row / column)
    maximum_path_each     DEX-TTS/model/monotonic_align/core.pyx:9-37      (forward DP over the band, then the backtrack)

Parity pin: the reference's own core.pyx, compiled in the build container from where it lies (oracle/make_golden_mas.py copies it to
/tmp and builds it with Cython -- ``oracle/_ref``-style, nothing of it enters the repo), run on seeded inputs ->
tests/golden/mas_*.npz; tests/test_mas_oracle.py replays them bit-exactly.
"""
import numpy as np

MAX_NEG = np.float32(-1e9)


def maximum_path_each(value, t_x, t_y):
    """value (Tx, Ty) float32, modified in place like upstream -> path (Tx, Ty) int32."""
    path = np.zeros(value.shape, dtype=np.int32)
    for y in range(t_y):                                                    # core.pyx:17-30, one column at a time
        lo, hi = max(0, t_x + y - t_y), min(t_x, y + 1)
        if hi <= lo:
            continue
        xs = np.arange(lo, hi)
        yp = max(y - 1, 0)                                                  # y = 0 only ever takes the constants below
        v_cur = np.where(xs == y, MAX_NEG, value[xs, yp])
        v_prev = np.where(xs == 0, np.float32(0.0) if y == 0 else MAX_NEG, value[np.maximum(xs - 1, 0), yp])
        value[xs, y] = np.maximum(v_cur, v_prev) + value[xs, y]
    index = t_x - 1
    for y in range(t_y - 1, -1, -1):                                        # core.pyx:32-35
        path[index, y] = 1
        if index != 0 and (index == y or value[index, y - 1] < value[index - 1, y - 1]):
            index -= 1
    return path


def maximum_path(value, mask):
    """value, mask (B, Tx, Ty) numpy -> path (B, Tx, Ty) float32 of zeros and ones."""
    value = (value * mask).astype(np.float32)
    t_xs = mask.sum(1)[:, 0].astype(np.int32)
    t_ys = mask.sum(2)[:, 0].astype(np.int32)
    return np.stack([maximum_path_each(value[b], int(t_xs[b]), int(t_ys[b])) for b in range(value.shape[0])]).astype(np.float32)
