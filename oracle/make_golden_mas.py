"""Generate tests/golden/mas_*.npz with the reference's own Cython kernel (DEX-TTS/model/monotonic_align/core.pyx), compiled from where
it lies: the .pyx is copied to a scratch directory under /tmp, built there with Cython (the shipped binary is stale) and driven through
the lines of ``maximum_path`` (DEX-TTS/model/monotonic_align/__init__.py:8-25).  Run in the build container only:
    python oracle/make_golden_mas.py
"""
import importlib.util
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(os.environ.get("DEX_REFERENCE_ROOT", "/root/reference"), "DEX-TTS", "model", "monotonic_align", "core.pyx")
SETUP = """from setuptools import setup, Extension
from Cython.Build import cythonize
import numpy
setup(name='mas_ref', ext_modules=cythonize([Extension('core', ['core.pyx'], include_dirs=[numpy.get_include()])], language_level=3))
"""
CASES = [("mas_b1", 1, 7, 19, 41), ("mas_b3r", 3, 40, 150, 42), ("mas_square", 2, 12, 12, 43)]       # name, B, Tx, Ty, seed


def build_reference():
    tmp = tempfile.mkdtemp(prefix="mas_ref_")
    shutil.copy(REF, os.path.join(tmp, "core.pyx"))
    with open(os.path.join(tmp, "setup.py"), "w") as f:
        f.write(SETUP)
    subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=tmp, check=True, capture_output=True)
    so = [n for n in os.listdir(tmp) if n.startswith("core") and n.endswith(".so")][0]
    spec = importlib.util.spec_from_file_location("core", os.path.join(tmp, so))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def synth_mas(B, Tx, Ty, seed):
    """A log-likelihood-like score matrix (negative, with a noisy diagonal ridge) and ragged (x, y) lengths; Ty >= Tx per utterance."""
    rng = np.random.default_rng(seed)
    value = (-3.0 * rng.random((B, Tx, Ty)) - 1.0).astype(np.float32)
    t_x = np.full(B, Tx)
    t_y = np.full(B, Ty)
    if B > 1:
        t_x[1:] = rng.integers(max(1, Tx // 2), Tx + 1, B - 1)
        t_y[1:] = np.maximum(t_x[1:], rng.integers(max(1, Ty // 2), Ty + 1, B - 1))
    for b in range(B):
        for x in range(t_x[b]):
            c = int((x + 0.5) * t_y[b] / t_x[b])
            value[b, x, max(0, c - 2):c + 3] += 2.0
    mask = ((np.arange(Tx)[None, :, None] < t_x[:, None, None]) & (np.arange(Ty)[None, None, :] < t_y[:, None, None])).astype(np.float32)
    return value, mask


def run_case(core, name, B, Tx, Ty, seed):
    value, mask = synth_mas(B, Tx, Ty, seed)
    # __init__.py:13-24, verbatim on numpy inputs
    v = (value * mask).astype(np.float32)
    path = np.zeros_like(v).astype(np.int32)
    t_x_max = mask.sum(1)[:, 0].astype(np.int32)
    t_y_max = mask.sum(2)[:, 0].astype(np.int32)
    core.maximum_path_c(path, v, t_x_max, t_y_max)
    out = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(out, path=np.packbits(path.astype(np.uint8), axis=-1), shape=np.array(path.shape, dtype=np.int64),
                        meta=np.array([B, Tx, Ty, seed], dtype=np.int64))
    print(f"{name}: path {path.shape} ones per utterance {path.sum((1, 2)).tolist()} (= t_y {t_y_max.tolist()}) -> {os.path.relpath(out, ROOT)}")


if __name__ == "__main__":
    core = build_reference()
    for c in CASES:
        run_case(core, *c)
