"""``TIVEncoder`` / ``TVEncoder`` -- the reference-speech encoders of DeXTTS with their forward passes on hand-written sm_100a CUDA.

Replace (same constructor arguments, same ``forward`` signature and return value, same ``state_dict`` keys):
    DEX-TTS/model/ref_encoder.py:83-107    class TIVEncoder  (attached as ``DeXTTS.tiv_encoder``, DEX-TTS/model/tts.py:28,50)
    DEX-TTS/model/ref_encoder.py:109-140   class TVEncoder   (attached as ``DeXTTS.tv_encoder``,  DEX-TTS/model/tts.py:26,43)
    DEX-TTS/model/ref_encoder.py:36-56     class LF0Encoder  (attached as ``DeXTTS.lf0_encoder``, DEX-TTS/model/tts.py:27,42)
    DEX-TTS/model/tts.py:45-49             the style fusion lines of ``DeXTTS.forward`` -> ``style_fusion(conv_sty, ...)``

This is the once-per-utterance stage right in front of the reverse-diffusion loop: the TIV encoder's six
skip tensors are the ``ref`` argument of ``Diffusion.forward``, the TV encoder's ``z_dec`` becomes its ``sty`` (tts.py:48-49).  Parameters and BatchNorm buffers are registered under the reference's names
(``in_conv.conv.weight``, ``in_conv.bn.running_mean``, ``conv_blocks.3.conv_block.1.conv.weight`` ...), so upstream checkpoints
load with ``strict=True``.  Inference (eval mode) only: BatchNorm uses its running statistics.
"""
import ctypes

import torch
import torch.nn as nn

from .. import lib as _lib
from ..synth import lf0_manifest, tiv_manifest, tv_manifest
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


def _register(root, manifest):
    """Rebuild a reference module tree (parameters and buffers under their upstream names) from a manifest."""
    for name, shape, kind in manifest:
        parts = name.split(".")
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, _Node())
            mod = mod._modules[p]
        if kind == "conv":                                  # nn.Conv1d default init: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
            bound = 1.0 / (shape[1] * shape[2]) ** 0.5
            mod.register_parameter(parts[-1], nn.Parameter(torch.empty(shape).uniform_(-bound, bound)))
        elif kind == "bias":
            mod.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape)))
        elif kind == "bn_w":
            mod.register_parameter(parts[-1], nn.Parameter(torch.ones(shape)))
        elif kind == "bn_b":
            mod.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape)))
        elif kind == "bn_rm":
            mod.register_buffer(parts[-1], torch.zeros(shape))
        elif kind == "bn_rv":
            mod.register_buffer(parts[-1], torch.ones(shape))
        elif kind == "code":                                # VQEmbeddingEMA: U(-1/n_emb, 1/n_emb) buffers (ref_encoder.py:192-197)
            mod.register_buffer(parts[-1], torch.empty(shape).uniform_(-1.0 / shape[0], 1.0 / shape[0]))
        elif kind == "gru":                                 # nn.GRU: U(-1/sqrt(hidden), 1/sqrt(hidden)), hidden = weight_hh.shape[1]
            mod.register_parameter(parts[-1], nn.Parameter(torch.empty(shape).uniform_(-1.0, 1.0) / (shape[0] // 3) ** 0.5))
        elif kind == "ema_n":
            mod.register_buffer(parts[-1], torch.zeros(shape))
        else:
            mod.register_buffer(parts[-1], torch.tensor(0, dtype=torch.long))


class TVEncoderEngine:
    """ctypes driver of the ``dexb_tv_*`` entry points (include/dexb200.h)."""

    SKIP = ("bn_n", "ema_n")                                # training bookkeeping the forward pass never reads

    def __init__(self, c_in, c_out, c_out_g, num_layer, c_h, n_emb, commit_w):
        if not torch.cuda.is_available():
            raise RuntimeError("dexb200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.dims = (int(c_in), int(c_out), int(c_out_g), int(num_layer), int(c_h), int(n_emb))
        self.L = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(self.L.dexb_tv_create(self.dims[0], self.dims[4], self.dims[1], self.dims[2], self.dims[3], self.dims[5],
                                         float(commit_w), ctypes.byref(h)), "dexb_tv_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.dexb_tv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_state_dict(self, sd, prefix="tv_encoder."):
        dev = torch.device("cuda", torch.cuda.current_device())
        for name, shape, kind in tv_manifest(*self.dims):
            if kind in self.SKIP or name == "vq.ema_weight":
                continue
            key = prefix + name
            if key not in sd:
                raise RuntimeError(f"state dict is missing '{key}'")
            t = sd[key].detach().to(device=dev, dtype=torch.float32).contiguous()
            if tuple(t.shape) != tuple(shape):
                raise RuntimeError(f"'{key}' has shape {tuple(t.shape)}, expected {tuple(shape)}")
            shp = (ctypes.c_int64 * t.dim())(*t.shape)
            _lib.check(self.L.dexb_tv_load_weight(self.h, name.encode(), ctypes.c_void_p(t.data_ptr()), shp, t.dim()),
                       f"dexb_tv_load_weight({name})")
        torch.cuda.synchronize()
        _lib.check(self.L.dexb_tv_finalize_weights(self.h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "dexb_tv_finalize_weights")

    def forward(self, sty, mask, return_indices=False):
        """sty (B, c_in, T) CUDA fp32, mask (B,1,T) or (B,T) -> (z_beforeVQ (B,c_out,T), z_dec (B,c_out_g,T), vq_loss 0-dim)
        [+ code indices (B,T) int32]."""
        c_in, c_out, c_out_g = self.dims[:3]
        B, C, T = sty.shape
        if C != c_in:
            raise RuntimeError(f"style features have {C} channels, the encoder expects {c_in}")
        sty = sty.detach().float().contiguous()
        m = mask.detach().float().reshape(B, T).contiguous()
        z_before = torch.empty(B, c_out, T, device=sty.device, dtype=torch.float32)
        z_dec = torch.empty(B, c_out_g, T, device=sty.device, dtype=torch.float32)
        loss = torch.empty((), device=sty.device, dtype=torch.float32)
        idx = torch.empty(B, T, device=sty.device, dtype=torch.int32) if return_indices else None
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        _lib.check(self.L.dexb_tv_forward(self.h, p(sty), p(m), B, T, p(z_before), p(z_dec), p(loss), p(idx) if return_indices else None,
                                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "dexb_tv_forward")
        self._keep = (sty, m)
        return (z_before, z_dec, loss, idx) if return_indices else (z_before, z_dec, loss)

    @property
    def launches(self):
        return int(self.L.dexb_tv_last_launch_count(self.h))


class TVEncoder(nn.Module):
    """DEX-TTS/model/ref_encoder.py:109-140."""

    def __init__(self, c_in, c_out, c_out_g, num_layer, c_h, n_emb, commit_w):
        super().__init__()
        self.dims = (int(c_in), int(c_out), int(c_out_g), int(num_layer), int(c_h), int(n_emb))
        self.commit_w = float(commit_w)
        _register(self, tv_manifest(*self.dims))
        self._engine = None
        self._sig = None

    def _signature(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def cuda_engine(self):
        sig = self._signature()
        if self._engine is None:
            self._engine = TVEncoderEngine(*self.dims, self.commit_w)
            self._sig = None
        if sig != self._sig:
            self._engine.load_state_dict(self.state_dict(), prefix="")
            self._sig = sig
        return self._engine

    @torch.no_grad()
    def forward(self, x, mask):
        if self.training:
            raise NotImplementedError("the CUDA TV encoder implements eval mode (no EMA codebook update, BatchNorm on running "
                                      "statistics) only")
        if not x.is_cuda:
            raise RuntimeError("dexb200.TVEncoder runs on CUDA (sm_100a) only; move the model and inputs to the GPU")
        if x.dim() == 4:
            x = x.squeeze(1)                                # ref_encoder.py:128
        return self.cuda_engine().forward(x, mask)


class LF0EncoderEngine:
    """ctypes driver of the ``dexb_lf0_*`` entry points (include/dexb200.h)."""

    def __init__(self, c_h, c_out, c_out_g, num_layer):
        if not torch.cuda.is_available():
            raise RuntimeError("dexb200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.dims = (int(c_h), int(c_out), int(c_out_g), int(num_layer))
        self.L = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(self.L.dexb_lf0_create(*self.dims, ctypes.byref(h)), "dexb_lf0_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.dexb_lf0_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_state_dict(self, sd, prefix="lf0_encoder."):
        dev = torch.device("cuda", torch.cuda.current_device())
        for name, shape, _ in lf0_manifest(*self.dims):
            key = prefix + name
            if key not in sd:
                raise RuntimeError(f"state dict is missing '{key}'")
            t = sd[key].detach().to(device=dev, dtype=torch.float32).contiguous()
            if tuple(t.shape) != tuple(shape):
                raise RuntimeError(f"'{key}' has shape {tuple(t.shape)}, expected {tuple(shape)}")
            shp = (ctypes.c_int64 * t.dim())(*t.shape)
            _lib.check(self.L.dexb_lf0_load_weight(self.h, name.encode(), ctypes.c_void_p(t.data_ptr()), shp, t.dim()),
                       f"dexb_lf0_load_weight({name})")
        torch.cuda.synchronize()
        _lib.check(self.L.dexb_lf0_finalize_weights(self.h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "dexb_lf0_finalize_weights")

    def forward(self, lf0, mask):
        """lf0 (B, T) CUDA fp32, mask (B,1,T) or (B,T) -> (lf0_enc (B,c_out,T), lf0_dec (B,c_out_g,T))."""
        B, T = lf0.shape
        lf0 = lf0.detach().float().contiguous()
        m = mask.detach().float().reshape(B, T).contiguous()
        enc = torch.empty(B, self.dims[1], T, device=lf0.device, dtype=torch.float32)
        dec = torch.empty(B, self.dims[2], T, device=lf0.device, dtype=torch.float32)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        _lib.check(self.L.dexb_lf0_forward(self.h, p(lf0), p(m), B, T, p(enc), p(dec),
                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "dexb_lf0_forward")
        self._keep = (lf0, m)
        return enc, dec

    @property
    def launches(self):
        return int(self.L.dexb_lf0_last_launch_count(self.h))


class LF0Encoder(nn.Module):
    """DEX-TTS/model/ref_encoder.py:36-56."""

    def __init__(self, c_h, c_out, c_out_g, num_layer, c_in=1):
        super().__init__()
        if c_in != 1:
            raise NotImplementedError("the CUDA LF0 encoder implements the one-channel contour input of the shipped configs")
        self.dims = (int(c_h), int(c_out), int(c_out_g), int(num_layer))
        _register(self, lf0_manifest(*self.dims))
        self._engine = None
        self._sig = None

    def _signature(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def cuda_engine(self):
        sig = self._signature()
        if self._engine is None:
            self._engine = LF0EncoderEngine(*self.dims)
            self._sig = None
        if sig != self._sig:
            self._engine.load_state_dict(self.state_dict(), prefix="")
            self._sig = sig
        return self._engine

    @torch.no_grad()
    def forward(self, lf0, mask):
        if not lf0.is_cuda:
            raise RuntimeError("dexb200.LF0Encoder runs on CUDA (sm_100a) only; move the model and inputs to the GPU")
        return self.cuda_engine().forward(lf0, mask)


@torch.no_grad()
def style_fusion(conv_sty, sty_enc, sty_dec, sty_mask, lf0_enc, lf0_dec, lf0_mask, want_sty_enc=True):
    """The four fusion lines of ``DeXTTS.forward`` (DEX-TTS/model/tts.py:45-49) as one C-ABI call (``dexb_style_fuse``):

        sty_enc = (sty_enc.sum(-1) / sty_mask.sum(-1)) + (lf0_enc.sum(-1) / lf0_mask.sum(-1));  sty_enc = sty_enc.squeeze(1)
        sty_dec = sty_dec + (lf0_dec.sum(-1) / lf0_mask.sum(-1)).unsqueeze(-1);                 sty_dec = self.conv_sty(sty_dec)

    ``conv_sty`` is the module's own ``nn.Conv1d(c_out_g, 2 * dim, 1)`` (tts.py:31; only its weight / bias tensors are read).
    -> (sty_enc (B, C) or None, sty_dec (B, 2*dim, Ts) = the ``sty`` argument of the decoder)."""
    if not sty_dec.is_cuda:
        raise RuntimeError("dexb200.style_fusion runs on CUDA (sm_100a) only")
    L = _lib.load()
    f = lambda t: t.detach().float().contiguous()
    B, C, Ts = sty_dec.shape
    Tl = lf0_dec.shape[-1]
    w, b = f(conv_sty.weight), f(conv_sty.bias)
    N = w.shape[0]
    if tuple(w.shape) != (N, C, 1):
        raise RuntimeError(f"conv_sty.weight has shape {tuple(w.shape)}, expected ({N}, {C}, 1)")
    z_before, z_dec, le, ld = f(sty_enc), f(sty_dec), f(lf0_enc), f(lf0_dec)
    # dexb_style_fuse takes ONE channel count for the four tensors (every shipped config has tv_encoder.c_out == c_out_g == the LF0
    # encoder's widths); anything else would read out of bounds inside the kernel
    for name, t, T_ in (("sty_enc", z_before, Ts), ("lf0_enc", le, Tl), ("lf0_dec", ld, Tl)):
        if tuple(t.shape) != (B, C, T_):
            raise RuntimeError(f"style_fusion: {name} has shape {tuple(t.shape)}, expected ({B}, {C}, {T_}) -- the TV / LF0 encoder widths "
                               f"(c_out, c_out_g) must all equal {C} for dexb_style_fuse")
    sm, lm = f(sty_mask).reshape(B, Ts), f(lf0_mask).reshape(B, Tl)
    scratch = torch.empty(B, C, device=z_dec.device, dtype=torch.float32)
    out_enc = torch.empty(B, C, device=z_dec.device, dtype=torch.float32) if want_sty_enc else None
    sty = torch.empty(B, N, Ts, device=z_dec.device, dtype=torch.float32)
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    _lib.check(L.dexb_style_fuse(p(z_before), p(z_dec), p(sm), Ts, p(le), p(ld), p(lm), Tl, B, C, p(w), p(b), N, p(scratch),
                                 p(out_enc), p(sty), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "dexb_style_fuse")
    return out_enc, sty
