"""``model.monotonic_align.maximum_path`` on the GPU (``dexb_mas_maximum_path``, csrc/mas.cu) -- replaces the reference's Cython
kernel and its device -> host -> device round trip (DEX-TTS/model/monotonic_align/__init__.py:8-25, core.pyx; call site
DEX-TTS/model/tts.py:108, training).  SURVEY.md §8f rank 4.  No CPU / PyTorch fallback."""
import ctypes

import torch

from .. import lib as _lib


@torch.no_grad()
def maximum_path(value, mask):
    """value, mask (B, Tx, Ty) on CUDA (mask = x_mask (x) y_mask in {0, 1}) -> path (B, Tx, Ty) of zeros and ones, value's dtype."""
    if not value.is_cuda:
        raise RuntimeError("dexb200.maximum_path runs on CUDA (sm_100a) only; move the inputs to the GPU")
    L = _lib.load()
    v = value.detach().float().contiguous()
    m = mask.detach().float().contiguous()
    B, Tx, Ty = v.shape
    scratch = torch.empty(B * Tx * Ty, dtype=torch.uint8, device=v.device)
    path = torch.empty(B, Tx, Ty, dtype=torch.float32, device=v.device)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(L.dexb_mas_maximum_path(p(v), p(m), B, Tx, Ty, p(scratch), p(path),
                                       ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "dexb_mas_maximum_path")
    return path.to(dtype=value.dtype)
