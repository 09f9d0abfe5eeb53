"""Mirror of the reference's ``model.utils`` for the names its callers import (``fix_len_compatibility``: DEX-TTS/main.py:14,
DEX-TTS/synthesize.py; ``sequence_mask``) plus the duration / alignment glue of ``DeXTTS.forward`` (DEX-TTS/model/tts.py:55-68,
GeDEX-TTS/model/tts.py:37-50) as two C-ABI calls (``dexb_align_lengths`` / ``dexb_align_expand``, csrc/align.cu).

``sequence_mask`` and ``fix_len_compatibility`` are host-side integer helpers (a comparison against ``arange`` and a loop over
one Python int); the alignment itself has no CPU / PyTorch fallback."""
import ctypes

import torch

from .. import lib as _lib


def sequence_mask(length, max_length=None):
    """DEX-TTS/model/utils.py:6-10: (B,) lengths -> (B, max_length) bool, ``t < length[b]``."""
    if max_length is None:
        max_length = length.max()
    t = torch.arange(int(max_length), dtype=length.dtype, device=length.device)
    return t.unsqueeze(0) < length.unsqueeze(1)


def fix_len_compatibility(length, num_downsamplings_in_unet=2):
    """DEX-TTS/model/utils.py:13-17: the next multiple of 2**num_downsamplings_in_unet (the reference counts up to it)."""
    q = 2 ** num_downsamplings_in_unet
    return -(-length // q) * q


_pinned = {}


@torch.no_grad()
def align_durations(logw, x_mask, mu_x, length_scale=1.0, want_attn=True):
    """tts.py:55-68 in two launches around the reference's own host round trip:

        w_ceil = ceil(exp(logw) * x_mask) * length_scale;  y_lengths = clamp_min(sum(w_ceil, [1, 2]), 1).long()
        y_max_length = int(y_lengths.max());  y_max_length_ = fix_len_compatibility(y_max_length)
        y_mask = sequence_mask(y_lengths, y_max_length_)[:, None];  attn = generate_path(w_ceil, x_mask (x) y_mask)[:, None]
        mu_y = (attn^T mu_x^T)^T

    logw, x_mask (B, 1, Tx), mu_x (B, n_feats, Tx) on CUDA -> (mu_y (B, n_feats, Ty_), y_mask (B, 1, Ty_), attn (B, 1, Tx, Ty_) or None,
    y_lengths (B,) int64 on the device, y_max_length: int).  The caller slices ``[:, :, :y_max_length]`` as tts.py:69-74 does."""
    if not logw.is_cuda:
        raise RuntimeError("dexb200.align_durations runs on CUDA (sm_100a) only; move the model and inputs to the GPU")
    L = _lib.load()
    f = lambda t: t.detach().float().contiguous()
    B, F, Tx = mu_x.shape
    lw, xm, mx = f(logw).reshape(B, Tx), f(x_mask).reshape(B, Tx), f(mu_x)
    dev = mx.device
    cum = torch.empty(B, Tx, device=dev, dtype=torch.float32)
    y_lengths = torch.empty(B, device=dev, dtype=torch.int64)
    host = _pinned.get(B)
    if host is None:
        host = _pinned[B] = torch.empty(B, dtype=torch.int64).pin_memory()
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.dexb_align_lengths(p(lw), p(xm), B, Tx, float(length_scale), p(cum), p(y_lengths), p(host), st), "dexb_align_lengths")
    y_max_length = int(host.max())
    Ty = int(fix_len_compatibility(y_max_length))
    attn = torch.empty(B, 1, Tx, Ty, device=dev, dtype=torch.float32) if want_attn else None
    y_mask = torch.empty(B, 1, Ty, device=dev, dtype=torch.float32)
    mu_y = torch.empty(B, F, Ty, device=dev, dtype=torch.float32)
    _lib.check(L.dexb_align_expand(p(cum), p(xm), p(y_lengths), p(mx), B, Tx, F, Ty, p(attn), p(y_mask), p(mu_y), st), "dexb_align_expand")
    return mu_y, y_mask, attn, y_lengths, y_max_length
