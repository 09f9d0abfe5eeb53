"""``TextEncoder`` -- the text encoder of DeXTTS / GeDEXTTS with its forward pass on hand-written sm_100a CUDA (``dexb_text_*``).

Replaces (same constructor arguments, same ``forward`` signature and return value, same ``state_dict`` keys):
    DEX-TTS/model/text_encoder.py:97-142     class TextEncoder  (attached as ``DeXTTS.encoder``, DEX-TTS/model/tts.py:29,51)
    GeDEX-TTS/model/text_encoder.py:99-146   class TextEncoder  (``GeDEXTTS.encoder``, GeDEX-TTS/model/tts.py:24,34) -> ``GeTextEncoder``

SURVEY.md §8f rank 2.  Eval mode (``n_spks > 1``: GeDEX-TTS only, as upstream) and the RetNet settings of the shipped configs (``use_softmax=True``,
``use_decay=False``; GLU feed-forward, pre-RMSNorm -- the defaults of model/retnet_cfg.py the reference never overrides).  Parameters
and the two RetNetRelPos buffers are registered under the reference's names (``emb.weight``, ``prenet.conv_layers.0.weight``,
``encoder.layers.3.retention.q_proj.weight``, ``encoder.retnet_rel_pos.angle``, ``proj_w.norm_1.gamma`` ...), so upstream checkpoints
load with ``strict=True``.  There is no CPU / PyTorch fallback.
"""
import ctypes
import math

import torch
import torch.nn as nn

from .. import lib as _lib
from ..synth import text_manifest
from .diffusion import _Node
from .utils import sequence_mask


class TextEncoderEngine:
    """ctypes driver of the ``dexb_text_*`` entry points (include/dexb200.h).  One handle = one (device, weights) pair."""

    def __init__(self, n_vocab, n_feats, n_channels, filter_channels, filter_channels_dp, n_heads, n_layers, kernel_size, adaln=True,
                 spk_emb_dim=0):
        if not torch.cuda.is_available():
            raise RuntimeError("dexb200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.dims = dict(n_vocab=int(n_vocab), n_feats=int(n_feats), n_channels=int(n_channels), filter_channels=int(filter_channels),
                         filter_channels_dp=int(filter_channels_dp), n_heads=int(n_heads), n_layers=int(n_layers),
                         kernel_size=int(kernel_size), adaln=bool(adaln), spk_emb_dim=int(spk_emb_dim))
        d = self.dims
        self.L = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(self.L.dexb_text_create(d["n_vocab"], d["n_feats"], d["n_channels"], d["filter_channels"], d["filter_channels_dp"],
                                           d["n_heads"], d["n_layers"], d["kernel_size"], int(d["adaln"]), d["spk_emb_dim"], ctypes.byref(h)),
                   "dexb_text_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.dexb_text_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_state_dict(self, sd, prefix="encoder."):
        dev = torch.device("cuda", torch.cuda.current_device())
        for name, shape, kind in text_manifest(**self.dims):
            if kind == "decay":                              # RetNetRelPos.decay: unused with use_decay=False
                continue
            key = prefix + name
            if key not in sd:
                raise RuntimeError(f"state dict is missing '{key}'")
            t = sd[key].detach().to(device=dev, dtype=torch.float32).contiguous()
            if tuple(t.shape) != tuple(shape):
                raise RuntimeError(f"'{key}' has shape {tuple(t.shape)}, expected {tuple(shape)}")
            shp = (ctypes.c_int64 * t.dim())(*t.shape)
            _lib.check(self.L.dexb_text_load_weight(self.h, name.encode(), ctypes.c_void_p(t.data_ptr()), shp, t.dim()),
                       f"dexb_text_load_weight({name})")
        torch.cuda.synchronize()
        _lib.check(self.L.dexb_text_finalize_weights(self.h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "dexb_text_finalize_weights")

    def forward(self, x, x_mask, sty=None, spk=None):
        """x (B, Tx) int64 CUDA, x_mask (B, 1, Tx) or (B, Tx), sty (B, C) or None, spk (B, spk_emb_dim) or None (n_spks > 1) ->
        (mu (B, n_feats, Tx), logw (B, 1, Tx))."""
        B, Tx = x.shape
        ids = x.detach().to(torch.int64).contiguous()
        m = x_mask.detach().float().reshape(B, Tx).contiguous()
        s = sty.detach().float().reshape(B, self.dims["n_channels"]).contiguous() if sty is not None else None
        if (spk is not None) != (self.dims["spk_emb_dim"] > 0):
            raise RuntimeError("the speaker embedding is " + ("required: this encoder was built for n_spks > 1" if spk is None
                                                              else "not taken by an encoder built for n_spks <= 1"))
        k = spk.detach().float().reshape(B, self.dims["spk_emb_dim"]).contiguous() if spk is not None else None
        mu = torch.empty(B, self.dims["n_feats"], Tx, device=x.device, dtype=torch.float32)
        logw = torch.empty(B, 1, Tx, device=x.device, dtype=torch.float32)
        p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        _lib.check(self.L.dexb_text_forward(self.h, p(ids), p(m), p(s), p(k), B, Tx, p(mu), p(logw),
                                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "dexb_text_forward")
        self._keep = (ids, m, s, k)
        return mu, logw

    def forward_stream(self, x, x_mask, sty=None, n_layers=0, spk=None):
        """Unit-parity aid: run the prenet and the first ``n_layers`` RetNet layers only -> the residual stream (B, Tx, C)."""
        _lib.check(self.L.dexb_text_set_layer_limit(self.h, int(n_layers)), "dexb_text_set_layer_limit")
        try:
            self.forward(x, x_mask, sty, spk)
            out = torch.empty(x.shape[0], x.shape[1], self.dims["n_channels"] + self.dims["spk_emb_dim"], device=x.device,
                              dtype=torch.float32)
            _lib.check(self.L.dexb_text_copy_stream(self.h, ctypes.c_void_p(out.data_ptr()),
                                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "dexb_text_copy_stream")
        finally:
            _lib.check(self.L.dexb_text_set_layer_limit(self.h, -1), "dexb_text_set_layer_limit")
        return out

    @property
    def launches(self):
        return int(self.L.dexb_text_last_launch_count(self.h))


def _register_text(root, manifest, n_channels):
    """Parameters / buffers under the reference's names.  Initial values: the tensors upstream zero-initialises stay zero (prenet.proj,
    text_encoder.py:52-53; the AdaLN W_scale / W_bias weights with biases 1 / 0, base.py:174-178), the embedding is N(0, C^-0.5)
    (:113-114), the two RetNetRelPos buffers are upstream's closed forms (retention.py:75-86); every other weight gets a plain
    U(-1, 1) / sqrt(fan_in) draw (upstream: xavier / nn.Linear defaults) -- real use loads a checkpoint over them."""
    n_heads = [shape[0] for _, shape, kind in manifest if kind == "decay"][0]
    for name, shape, kind in manifest:
        parts = name.split(".")
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, _Node())
            mod = mod._modules[p]
        leaf = parts[-1]
        if kind == "emb":
            t = torch.empty(shape).normal_(0.0, n_channels ** -0.5)
        elif kind == "conv":
            t = torch.zeros(shape) if name.startswith("prenet.proj") else \
                torch.empty(shape).uniform_(-1, 1) / (shape[1] * shape[2]) ** 0.5
        elif kind == "lin":
            t = torch.empty(shape).uniform_(-1, 1) / shape[1] ** 0.5
        elif kind == "ada":
            t = torch.zeros(shape)
        elif kind == "bias":
            t = torch.zeros(shape)
        elif kind == "bn_w":                                 # norm gains and the AdaLN W_scale bias: 1
            t = torch.ones(shape)
        elif kind == "bn_b":
            t = torch.zeros(shape)
        elif kind == "angle":
            a = 1.0 / (10000 ** torch.linspace(0, 1, shape[0] // 2))
            mod.register_buffer(leaf, a.unsqueeze(-1).repeat(1, 2).flatten())
            continue
        elif kind == "decay":
            mod.register_buffer(leaf, torch.log(1 - 2 ** (-5 - torch.arange(n_heads, dtype=torch.float))))
            continue
        else:
            raise ValueError(kind)
        mod.register_parameter(leaf, nn.Parameter(t))


class _TextEncoderBase(nn.Module):
    adaln = True

    def __init__(self, n_vocab, n_feats, n_channels, filter_channels, filter_channels_dp, n_heads, n_layers, kernel_size, p_dropout,
                 use_softmax, use_decay, window_size=None, spk_emb_dim=64, n_spks=1):
        super().__init__()
        if n_spks > 1 and self.adaln:
            raise NotImplementedError("n_spks > 1 exists for GeDEX-TTS only: DeXTTS forces n_spks = 0 (DEX-TTS/model/tts.py:18) and passes "
                                      "spk=None to its encoder (:52)")
        if not use_softmax or use_decay:
            raise NotImplementedError("the CUDA text encoder implements the shipped RetNet settings: use_softmax=True, use_decay=False")
        self.n_vocab, self.n_feats, self.n_channels, self.n_spks = n_vocab, n_feats, n_channels, n_spks
        self.dims = dict(n_vocab=int(n_vocab), n_feats=int(n_feats), n_channels=int(n_channels), filter_channels=int(filter_channels),
                         filter_channels_dp=int(filter_channels_dp), n_heads=int(n_heads), n_layers=int(n_layers),
                         kernel_size=int(kernel_size), adaln=self.adaln, spk_emb_dim=int(spk_emb_dim) if n_spks > 1 else 0)
        _register_text(self, text_manifest(**self.dims), n_channels)
        self._engine = None
        self._sig = None

    def _signature(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def cuda_engine(self):
        sig = self._signature()
        if self._engine is None:
            self._engine = TextEncoderEngine(**self.dims)
            self._sig = None
        if sig != self._sig:
            self._engine.load_state_dict(self.state_dict(), prefix="")
            self._sig = sig
        return self._engine

    def _run(self, x, x_lengths, sty, spk=None):
        if not x.is_cuda:
            raise RuntimeError("dexb200.TextEncoder runs on CUDA (sm_100a) only; move the model and inputs to the GPU")
        x_mask = torch.unsqueeze(sequence_mask(x_lengths, x.size(1)), 1).to(torch.float32)          # text_encoder.py:132
        mu, logw = self.cuda_engine().forward(x, x_mask, sty, spk if self.n_spks > 1 else None)
        return mu, logw, x_mask


class TextEncoder(_TextEncoderBase):
    """DEX-TTS/model/text_encoder.py:97-142."""
    adaln = True

    @torch.no_grad()
    def forward(self, x, x_lengths, sty, spk=None):
        return self._run(x, x_lengths, sty)


class GeTextEncoder(_TextEncoderBase):
    """GeDEX-TTS/model/text_encoder.py:99-146 (no AdaptiveLayerNorm, no style input)."""
    adaln = False

    @torch.no_grad()
    def forward(self, x, x_lengths, spk=None):
        return self._run(x, x_lengths, None, spk)                 # n_spks > 1: spk (B, spk_emb_dim) joins behind the prenet (:141-142)
