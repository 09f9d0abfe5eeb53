"""``TIVEncoder`` -- the time-invariant reference encoder of DeXTTS with its forward pass on hand-written sm_100a CUDA.

Replaces (same constructor arguments, same ``forward`` signature and return value, same ``state_dict`` keys):
    DEX-TTS/model/ref_encoder.py:83-107   class TIVEncoder   (attached as ``DeXTTS.tiv_encoder``, DEX-TTS/model/tts.py:28,50)

This is the once-per-utterance stage right in front of the reverse-diffusion loop: its six skip tensors
are the ``ref`` argument of ``Diffusion.forward``.  Parameters and BatchNorm buffers are registered under the reference's names
(``in_conv.conv.weight``, ``in_conv.bn.running_mean``, ``conv_blocks.3.conv_block.1.conv.weight`` ...), so upstream checkpoints
load with ``strict=True``.  Inference (eval mode) only: BatchNorm uses its running statistics.
"""
import ctypes

import torch
import torch.nn as nn

from .. import lib as _lib
from ..synth import tiv_manifest
from .diffusion import _Node


class TIVEncoderEngine:
    """ctypes driver of the ``dexb_tiv_*`` entry points (include/dexb200.h).  One handle = one (device, weights) pair."""

    def __init__(self, c_in, c_out, num_layer, c_h):
        if not torch.cuda.is_available():
            raise RuntimeError("dexb200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.dims = (int(c_in), int(c_out), int(num_layer), int(c_h))
        self.L = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(self.L.dexb_tiv_create(self.dims[0], self.dims[3], self.dims[1], self.dims[2], ctypes.byref(h)), "dexb_tiv_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.dexb_tiv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_state_dict(self, sd, prefix="tiv_encoder."):
        """Copy every float tensor of the reference TIVEncoder state dict (``prefix + name``) to the handle and pack it."""
        dev = torch.device("cuda", torch.cuda.current_device())
        c_in, c_out, num_layer, c_h = self.dims
        for name, shape, kind in tiv_manifest(c_in, c_out, num_layer, c_h):
            if kind == "bn_n":                              # num_batches_tracked: bookkeeping of training only
                continue
            key = prefix + name
            if key not in sd:
                raise RuntimeError(f"state dict is missing '{key}'")
            t = sd[key].detach().to(device=dev, dtype=torch.float32).contiguous()
            if tuple(t.shape) != tuple(shape):
                raise RuntimeError(f"'{key}' has shape {tuple(t.shape)}, expected {tuple(shape)}")
            shp = (ctypes.c_int64 * t.dim())(*t.shape)
            _lib.check(self.L.dexb_tiv_load_weight(self.h, name.encode(), ctypes.c_void_p(t.data_ptr()), shp, t.dim()),
                       f"dexb_tiv_load_weight({name})")
        torch.cuda.synchronize()
        _lib.check(self.L.dexb_tiv_finalize_weights(self.h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "dexb_tiv_finalize_weights")

    def forward(self, ref, mask, want_out=True):
        """ref (B, c_in, T) CUDA fp32, mask (B, 1, T) or (B, T) -> (out (B, c_out, T) or None, [num_layer x (B, c_h, T)])."""
        c_in, c_out, num_layer, c_h = self.dims
        B, C, T = ref.shape
        if C != c_in:
            raise RuntimeError(f"reference features have {C} channels, the encoder expects {c_in}")
        ref = ref.detach().float().contiguous()
        m = mask.detach().float().reshape(B, T).contiguous()
        skips = [torch.empty(B, c_h, T, device=ref.device, dtype=torch.float32) for _ in range(num_layer)]
        out = torch.empty(B, c_out, T, device=ref.device, dtype=torch.float32) if want_out else None
        arr = (ctypes.c_void_p * num_layer)(*[s.data_ptr() for s in skips])
        _lib.check(self.L.dexb_tiv_forward(self.h, ctypes.c_void_p(ref.data_ptr()), ctypes.c_void_p(m.data_ptr()), B, T,
                                           ctypes.c_void_p(out.data_ptr()) if want_out else None, arr,
                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "dexb_tiv_forward")
        self._keep = (ref, m)                               # inputs stay alive until the stream has consumed them
        return out, skips

    @property
    def launches(self):
        return int(self.L.dexb_tiv_last_launch_count(self.h))


class TIVEncoder(nn.Module):
    """DEX-TTS/model/ref_encoder.py:83-107."""

    def __init__(self, c_in, c_out, num_layer, c_h):
        super().__init__()
        self.dims = (int(c_in), int(c_out), int(num_layer), int(c_h))
        for name, shape, kind in tiv_manifest(*self.dims[:3], c_h=self.dims[3]):
            parts = name.split(".")
            mod = self
            for p in parts[:-1]:
                if p not in mod._modules:
                    mod.add_module(p, _Node())
                mod = mod._modules[p]
            if kind == "conv":                              # nn.Conv1d default init (kaiming_uniform, a = sqrt(5))
                bound = 1.0 / (shape[1] * shape[2]) ** 0.5
                mod.register_parameter(parts[-1], nn.Parameter(torch.empty(shape).uniform_(-bound, bound)))
            elif kind == "bn_w":
                mod.register_parameter(parts[-1], nn.Parameter(torch.ones(shape)))
            elif kind == "bn_b":
                mod.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape)))
            elif kind == "bn_rm":
                mod.register_buffer(parts[-1], torch.zeros(shape))
            elif kind == "bn_rv":
                mod.register_buffer(parts[-1], torch.ones(shape))
            else:
                mod.register_buffer(parts[-1], torch.tensor(0, dtype=torch.long))
        self._engine = None
        self._sig = None

    def _signature(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def cuda_engine(self):
        """The libdexb200 handle for the current tensors (re-packed whenever a parameter or buffer changed)."""
        sig = self._signature()
        if self._engine is None:
            self._engine = TIVEncoderEngine(*self.dims)
            self._sig = None
        if sig != self._sig:
            self._engine.load_state_dict(self.state_dict(), prefix="")
            self._sig = sig
        return self._engine

    @torch.no_grad()
    def forward(self, x, mask):
        if self.training:
            raise NotImplementedError("the CUDA TIV encoder implements eval mode (BatchNorm on running statistics) only")
        if not x.is_cuda:
            raise RuntimeError("dexb200.TIVEncoder runs on CUDA (sm_100a) only; move the model and inputs to the GPU")
        if x.dim() == 4:
            x = x.squeeze(1)                                # ref_encoder.py:97
        return self.cuda_engine().forward(x, mask)
