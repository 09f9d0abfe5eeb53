"""Generate tests/golden/align_*.npz: the duration / alignment glue of DeXTTS.forward (DEX-TTS/model/tts.py:55-68) executed
verbatim on the UNMODIFIED reference helpers ``model.utils.sequence_mask`` / ``fix_len_compatibility`` / ``generate_path``
(DEX-TTS/model/utils.py:6-39) over seeded text-encoder outputs.  Run in the build container only:
    python oracle/make_golden_align.py
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "dex-tts_b200"))

import ref_loader                                     # noqa: E402
from dexb200.synth import synth_align_inputs          # noqa: E402

CASES = [
    # name,         B, Tx,  ragged, seed, length_scale, mean_dur
    ("align_b1",    1, 23,  False,  61,   1.0,          3.0),
    ("align_b3r",   3, 128, True,   62,   1.0,          4.0),     # the text length of BASELINE.json's C2 / C3
    ("align_b2s",   2, 57,  True,   63,   1.25,         2.0),     # --length_scale (synthesize.py:126), dyadic so sums are exact
]


def run_case(name, B, Tx, ragged, seed, length_scale, mean_dur):
    ref_loader.load_reference("dex")
    U = importlib.import_module("model.utils")
    sequence_mask, fix_len_compatibility, generate_path = U.sequence_mask, U.fix_len_compatibility, U.generate_path
    inp = synth_align_inputs(B, Tx, seed=seed, ragged=ragged, mean_dur=mean_dur)
    mu_x, logw, x_mask = inp["mu_x"], inp["logw"], inp["x_mask"]
    with torch.no_grad():
        # tts.py:55-68, verbatim
        w             = torch.exp(logw) * x_mask
        w_ceil        = torch.ceil(w) * length_scale
        y_lengths     = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
        y_max_length  = int(y_lengths.max())
        y_max_length_ = fix_len_compatibility(y_max_length)
        y_mask    = sequence_mask(y_lengths, y_max_length_).unsqueeze(1).to(x_mask.dtype)
        attn_mask = x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)
        attn      = generate_path(w_ceil.squeeze(1), attn_mask.squeeze(1)).unsqueeze(1)
        mu_y    = torch.matmul(attn.squeeze(1).transpose(1, 2), mu_x.transpose(1, 2))
        mu_y    = mu_y.transpose(1, 2)
    arrs = dict(mu_y=mu_y.numpy(), y_mask=y_mask.numpy(), attn=np.packbits(attn.numpy().astype(np.uint8), axis=-1),
                attn_shape=np.array(attn.shape, dtype=np.int64), y_lengths=y_lengths.numpy(),
                meta=np.array([B, Tx, int(ragged), seed, y_max_length, y_max_length_], dtype=np.int64),
                scale=np.array([length_scale, mean_dur], dtype=np.float64))
    assert set(np.unique(attn.numpy())) <= {0.0, 1.0}
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: y_lengths {y_lengths.tolist()} Ty_ {y_max_length_} mu_y {tuple(mu_y.shape)} attn {tuple(attn.shape)} -> "
          f"{os.path.relpath(path, ROOT)} ({os.path.getsize(path)/1024:.0f} KiB)")


if __name__ == "__main__":
    for c in CASES:
        run_case(*c)
