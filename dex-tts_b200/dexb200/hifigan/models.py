"""HiFi-GAN v1 generator on the CUDA path (``dexb_voc_*``, csrc/vocoder.cu) behind the reference's own class:

    hifigan.Generator(h)                 DEX-TTS/hifigan/models.py:112-173   (built by get_vocoder, DEX-TTS/src/utils.py:251-281)
        .load_state_dict(ckpt["generator"])     weight-normed checkpoints: ``*.weight_g`` / ``*.weight_v`` / ``*.bias``
        .eval(); .remove_weight_norm(); .to(device)
        vocoder(y_dec) -> (B, 1, 256 T)         DEX-TTS/synthesize.py:106

The module is a parameter container with the reference's names and shapes (so checkpoints load unchanged, before or after
``remove_weight_norm``); ``forward`` runs every convolution on the tcgen05 implicit-GEMM engine.  There is no PyTorch / CPU fallback.
"""
import ctypes

import torch
import torch.nn as nn

from .. import lib as _lib

LRELU_SLOPE = 0.1


def vocoder_manifest(h):
    """[(name, shape)] of the generator AFTER remove_weight_norm (the names ``dexb_voc_load_weight`` takes)."""
    ch0 = int(h["upsample_initial_channel"])
    out = [("conv_pre.weight", (ch0, 80, 7)), ("conv_pre.bias", (ch0,))]
    for i, k in enumerate(h["upsample_kernel_sizes"]):
        out += [(f"ups.{i}.weight", (ch0 >> i, ch0 >> (i + 1), int(k))), (f"ups.{i}.bias", (ch0 >> (i + 1),))]
    nk = len(h["resblock_kernel_sizes"])
    for i in range(len(h["upsample_rates"])):
        ch = ch0 >> (i + 1)
        for j, k in enumerate(h["resblock_kernel_sizes"]):
            for grp in ("convs1", "convs2"):
                for d in range(len(h["resblock_dilation_sizes"][j])):
                    p = f"resblocks.{i * nk + j}.{grp}.{d}"
                    out += [(p + ".weight", (ch, ch, int(k))), (p + ".bias", (ch,))]
    ch = ch0 >> len(h["upsample_rates"])
    out += [("conv_post.weight", (1, ch, 7)), ("conv_post.bias", (1,))]
    return out


class VocoderEngine:
    """ctypes driver of the ``dexb_voc_*`` entry points (include/dexb200.h).  One handle = one (device, weights) pair."""

    def __init__(self, h):
        if not torch.cuda.is_available():
            raise RuntimeError("dexb200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.cfg = dict(h)
        rates = [int(u) for u in h["upsample_rates"]]
        kern = [int(k) for k in h["upsample_kernel_sizes"]]
        if any(k != 2 * u for k, u in zip(kern, rates)):
            raise RuntimeError(f"upsample kernels {kern} must be twice the rates {rates} (hifigan/config.json pairs them so)")
        if str(h.get("resblock", "1")) != "1":
            raise RuntimeError("only the v1 ResBlock (resblock = '1') is built")
        dil = [tuple(int(x) for x in d) for d in h["resblock_dilation_sizes"]]
        if any(d != dil[0] for d in dil):
            raise RuntimeError("the ResBlocks must share one dilation tuple (hifigan/config.json: (1, 3, 5) x 3)")
        rk = [int(k) for k in h["resblock_kernel_sizes"]]
        self.up = 1
        for u in rates:
            self.up *= u
        self.L = _lib.load()
        hd = ctypes.c_void_p()
        I = ctypes.c_int32
        _lib.check(self.L.dexb_voc_create(80, int(h["upsample_initial_channel"]), (I * len(rates))(*rates), len(rates),
                                          (I * len(rk))(*rk), len(rk), (I * len(dil[0]))(*dil[0]), len(dil[0]), ctypes.byref(hd)),
                   "dexb_voc_create")
        self.h = hd

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.dexb_voc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_state_dict(self, sd):
        """``sd``: the generator's state dict after remove_weight_norm (plain ``weight`` / ``bias`` per convolution)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        for name, shape in vocoder_manifest(self.cfg):
            if name not in sd:
                raise RuntimeError(f"state dict is missing '{name}'")
            t = sd[name].detach().to(device=dev, dtype=torch.float32).contiguous()
            if tuple(t.shape) != tuple(shape):
                raise RuntimeError(f"'{name}' has shape {tuple(t.shape)}, expected {tuple(shape)}")
            shp = (ctypes.c_int64 * t.dim())(*t.shape)
            _lib.check(self.L.dexb_voc_load_weight(self.h, name.encode(), ctypes.c_void_p(t.data_ptr()), shp, t.dim()),
                       f"dexb_voc_load_weight({name})")
        torch.cuda.synchronize()
        _lib.check(self.L.dexb_voc_finalize_weights(self.h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "dexb_voc_finalize_weights")

    def forward(self, mel):
        """mel (B, 80, T) CUDA fp32 -> waveform (B, 1, T * prod(upsample_rates))."""
        B, C, T = mel.shape
        if C != 80:
            raise RuntimeError(f"the vocoder expects 80 mel bins, got {C}")
        mel = mel.detach().float().contiguous()
        wav = torch.empty(B, 1, T * self.up, device=mel.device, dtype=torch.float32)
        with torch.cuda.device(mel.device):
            _lib.check(self.L.dexb_voc_forward(self.h, ctypes.c_void_p(mel.data_ptr()), B, T, ctypes.c_void_p(wav.data_ptr()),
                                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "dexb_voc_forward")
        self._keep = mel
        return wav

    @property
    def launches(self):
        return int(self.L.dexb_voc_last_launch_count(self.h))


class _WN(nn.Module):
    """One weight-normed convolution of the reference as a parameter holder: ``weight_g`` / ``weight_v`` / ``bias`` until
    ``fold()`` (= torch's remove_weight_norm: w = v * g / ||v||, norm over all dims but 0), then ``weight`` / ``bias``."""

    def __init__(self, shape, std=None):
        super().__init__()
        fan_in = shape[1] * shape[2]
        v = torch.empty(shape)
        if std is None:                                     # Conv1d default init (convs of conv_pre keep it upstream)
            bound = 1.0 / fan_in ** 0.5
            v.uniform_(-bound, bound)
        else:
            v.normal_(0.0, std)                             # init_weights, models.py:10-13
        bound = 1.0 / fan_in ** 0.5
        self.bias = nn.Parameter(torch.empty(self._n_bias(shape)).uniform_(-bound, bound))      # upstream key order: bias, g, v
        self.weight_g = nn.Parameter(torch.norm_except_dim(v, 2, 0))
        self.weight_v = nn.Parameter(v)

    @staticmethod
    def _n_bias(shape):
        return shape[0]

    def fold(self):
        if "weight" in self._parameters:
            return
        with torch.no_grad():
            w = torch._weight_norm(self.weight_v, self.weight_g, 0)     # what WeightNorm.remove computes
        del self._parameters["weight_g"], self._parameters["weight_v"]
        bias = self._parameters.pop("bias")                 # remove_weight_norm re-registers weight AFTER bias
        self.bias = bias
        self.weight = nn.Parameter(w)


class _WNT(_WN):
    """ConvTranspose1d: weight (in, out, k), bias (out,); weight_norm's default dim 0 is the INPUT channel there, as upstream."""

    @staticmethod
    def _n_bias(shape):
        return shape[1]


class _Res(nn.Module):
    def __init__(self, ch, k, n_dil):
        super().__init__()
        self.convs1 = nn.ModuleList([_WN((ch, ch, k), std=0.01) for _ in range(n_dil)])
        self.convs2 = nn.ModuleList([_WN((ch, ch, k), std=0.01) for _ in range(n_dil)])


class Generator(nn.Module):
    """DEX-TTS/hifigan/models.py:112-173.  ``h``: hifigan/config.json as an AttrDict (or any mapping)."""

    def __init__(self, h):
        super().__init__()
        self.h = h
        cfg = dict(h)
        self.num_kernels = len(cfg["resblock_kernel_sizes"])
        self.num_upsamples = len(cfg["upsample_rates"])
        ch0 = int(cfg["upsample_initial_channel"])
        self.conv_pre = _WN((ch0, 80, 7))
        self.ups = nn.ModuleList([_WNT((ch0 >> i, ch0 >> (i + 1), int(k)), std=0.01)
                                  for i, k in enumerate(cfg["upsample_kernel_sizes"])])
        self.resblocks = nn.ModuleList()
        for i in range(self.num_upsamples):
            for k, d in zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"]):
                self.resblocks.append(_Res(ch0 >> (i + 1), int(k), len(d)))
        self.conv_post = _WN((1, ch0 >> self.num_upsamples, 7), std=0.01)
        self._cfg = cfg
        self._engine = None
        self._sig = None
        self._folded = False

    def remove_weight_norm(self):
        print("Removing weight norm...")
        for m in self.modules():
            if isinstance(m, _WN):
                m.fold()
        self._folded = True

    def _signature(self):
        return tuple((t.data_ptr(), t._version) for t in self.parameters())

    def cuda_engine(self):
        sig = self._signature()
        if self._engine is None:
            self._engine = VocoderEngine(self._cfg)
            self._sig = None
        if sig != self._sig:
            self._engine.load_state_dict(self.state_dict())
            self._sig = sig
        return self._engine

    @torch.no_grad()
    def forward(self, x):
        if not self._folded:
            raise NotImplementedError("the CUDA generator runs the inference state of get_vocoder: call remove_weight_norm() first "
                                      "(DEX-TTS/src/utils.py:278)")
        if not x.is_cuda:
            raise RuntimeError("dexb200.hifigan.Generator runs on CUDA (sm_100a) only; move the vocoder and the mel to the GPU")
        return self.cuda_engine().forward(x)
