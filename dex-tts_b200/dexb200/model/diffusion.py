"""``Diffusion`` -- the decoder module of DeXTTS / GeDEXTTS with the reverse-diffusion loop on hand-written sm_100a CUDA.

Replaces (same constructor arguments, same ``forward`` signature, same ``state_dict`` keys):
    DEX-TTS/model/diffusion.py:238-259    class Diffusion   (attached as ``DeXTTS.decoder``,  DEX-TTS/model/tts.py:30)
    GeDEX-TTS/model/diffusion.py:209-229  class Diffusion   (attached as ``GeDEXTTS.decoder``, GeDEX-TTS/model/tts.py:25)

``forward(..., infer=True)`` draws the initial noise exactly like the reference (``torch.randn(shape, device) /
temperature + mu``, diffusion.py:256-257) and hands the trajectory to libdexb200.so (ablation_sampler + EDMPrecond +
DiffusionDenoiser, all steps in one CUDA graph).  Parameters stay ordinary ``nn.Parameter``s registered under the
reference's names -- including the aliasing of every denoiser tensor under both ``denoise_fn.*`` and
``precond_model.model.*`` (diffusion.py:242-243) -- so upstream checkpoints load with ``strict=True`` and ``.to()``,
EMA copies etc. keep working.  The training branch (``infer=False`` -> EDMLoss, diffusion.py:252-254) is not a CUDA path: it
delegates to the reference's own ``model.diffusion.Diffusion`` running on THIS module's parameters (``reference_twin.py``).
"""
import math

import torch
import torch.nn as nn

from ..engine import ReverseDiffusion
from ..manifest import DecoderCfg, decoder_manifest


class _Node(nn.Module):
    """Anonymous container used to rebuild the reference's module tree from dotted parameter names."""


def _build_tree(root, manifest, cfg):
    for e in manifest:
        parts = e.name.split(".")
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, _Node())
            mod = mod._modules[p]
        bound = 1.0 / math.sqrt(max(e.fan_in, 1))
        if e.init in ("conv", "lin"):
            t = torch.empty(e.shape).uniform_(-bound, bound)
        elif e.init == "one":
            t = torch.ones(e.shape)
        elif e.init == "zero":
            t = torch.zeros(e.shape)
        elif e.init == "posconv":                        # make_conv_pos, dit.py:84-85
            t = torch.zeros(e.shape) if e.name.endswith("bias") else \
                torch.empty(e.shape).normal_(0.0, math.sqrt(4.0 / (e.shape[2] * cfg.hidden)))
        else:
            raise ValueError(e.init)
        mod.register_parameter(parts[-1], nn.Parameter(t))


class _Precond(nn.Module):
    """Holds the denoiser a second time as ``.model`` (EDMPrecond, edm.py:74-86) so the state-dict keys match."""

    def __init__(self, model):
        super().__init__()
        self.model = model
        self.sigma_min, self.sigma_max, self.sigma_data = 0, float("inf"), 0.5


class _DiffusionBase(nn.Module):
    variant = None

    def __init__(self, n_feats, dim, dit_cfg, loss_type="base", precond="edm", model_type="vit", dim_mults=(1, 2), n_spks=1,
                 spk_emb_dim=64, pe_scale=1000, gemm_engine=0, nsplit=3):
        super().__init__()
        if tuple(dim_mults) != (1, 2):
            raise NotImplementedError("the CUDA path implements the two-level U-Net of the shipped configs (dim_mults=[1, 2])")
        if model_type not in ("dit", "vit"):
            raise NotImplementedError(f"model_type {model_type!r}")
        get = (lambda k: dit_cfg[k]) if isinstance(dit_cfg, dict) else (lambda k: getattr(dit_cfg, k))
        self.cfg = DecoderCfg.make(self.variant, dim=dim, hidden=get("hidden_size"), depth=get("depth"), heads=get("num_heads"),
                                   mlp_ratio=get("mlp_ratio"), patch=get("patch_size"), stride=get("stride_size"),
                                   conv_pos=get("conv_pos"), conv_pos_groups=get("conv_pos_groups"), n_feats=n_feats,
                                   pe_scale=pe_scale, n_spks=n_spks, spk_emb_dim=spk_emb_dim)
        self.denoise_fn = _Node()
        _build_tree(self.denoise_fn, decoder_manifest(self.cfg), self.cfg)
        self.precond_model = _Precond(self.denoise_fn)
        self._gemm_engine, self._nsplit = gemm_engine, nsplit
        self._engine = None
        self._sig = None
        self._twin = None
        self._ctor = dict(n_feats=n_feats, dim=dim, dit_cfg=dit_cfg, loss_type=loss_type, precond=precond, model_type=model_type,
                          dim_mults=tuple(dim_mults), n_spks=n_spks, spk_emb_dim=spk_emb_dim, pe_scale=pe_scale)

    def _training_twin(self):
        """The reference's own ``Diffusion`` on this module's parameters (training branch only, see reference_twin.py)."""
        if self._twin is None:
            from .reference_twin import diffusion_twin
            object.__setattr__(self, "_twin", diffusion_twin(self, self._ctor))     # not a registered sub-module: no duplicate keys
        self._twin.train(self.training)
        return self._twin

    # ---- CUDA engine management ---------------------------------------------------------------------
    def _weights_signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.denoise_fn.parameters())

    def cuda_engine(self):
        """The libdexb200 handle for the current weights (re-packed whenever a parameter tensor changed)."""
        sig = self._weights_signature()
        if self._engine is None:
            self._engine = ReverseDiffusion(self.cfg, gemm_engine=self._gemm_engine, nsplit=self._nsplit)
            self._sig = None
        if sig != self._sig:
            sd = {"denoise_fn." + k: v for k, v in self.denoise_fn.state_dict().items()}
            self._engine.load_state_dict(sd)
            self._sig = sig
        return self._engine

    def _run(self, x, mask, mu, n_timesteps, temperature, cond):
        if not mu.is_cuda:
            raise RuntimeError("dexb200.Diffusion runs on CUDA (sm_100a) only; move the model and inputs to the GPU")
        shape = (mu.shape[0], 80, mu.shape[2])                              # diffusion.py:256 (hard-coded 80 bins)
        x = torch.randn(shape, device=x.device) / temperature + mu
        return self.cuda_engine().sample(x, mask, mu, int(n_timesteps), cond=cond)


class Diffusion(_DiffusionBase):
    """DEX-TTS decoder (TV / TIV adaptors).  DEX-TTS/model/diffusion.py:238-259."""
    variant = "dex"

    def forward(self, x, mask, mu, ref, ref_lengths, sty, sty_lengths, n_timesteps=1, spk=None, infer=False, temperature=1.0,
                mask_ratio=0):
        if not infer:                                                       # EDMLoss: the reference's PyTorch, our parameters
            return self._training_twin()(x, mask, mu, ref, ref_lengths, sty, sty_lengths, n_timesteps=n_timesteps, spk=spk,
                                         infer=False, temperature=temperature, mask_ratio=mask_ratio)
        with torch.no_grad():
            cond = dict(sty=sty, sty_lengths=sty_lengths, ref_skips=list(ref))
            return self._run(x, mask, mu, n_timesteps, temperature, cond)


class GeDiffusion(_DiffusionBase):
    """GeDEX-TTS decoder (no reference speech).  GeDEX-TTS/model/diffusion.py:209-229."""
    variant = "gedex"

    def forward(self, x, mask, mu, n_timesteps=1, spk=None, infer=False, temperature=1.0, mask_ratio=0):
        if not infer:
            return self._training_twin()(x, mask, mu, n_timesteps=n_timesteps, spk=spk, infer=False, temperature=temperature,
                                         mask_ratio=mask_ratio)
        with torch.no_grad():
            # n_spks > 1: `spk` is the speaker embedding GeDEXTTS.forward looked up (tts.py:30-31); it becomes the third input channel
            cond = dict(spk=spk) if self.cfg.n_spks > 1 else None
            return self._run(x, mask, mu, n_timesteps, temperature, cond)
