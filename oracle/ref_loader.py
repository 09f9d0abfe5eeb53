"""Import the UNMODIFIED reference decoder (read-only, from /root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` to pin the oracle restatement against
the real reference.  The reference cannot travel to the GPU box, so nothing under ``tests/ -m gpu``,
``bench.py`` or ``__graft_entry__.smoke()`` imports this module.

Shims (none of them edits a reference file):
  * ``timm`` is not installed -> a stub package that restates timm's published ``Attention`` and ``Mlp``
    modules (qkv Linear -> (B,N,3,H,hd) -> softmax(q k^T * hd^-0.5) v -> proj; fc1 -> act -> fc2).
    Call sites: DEX-TTS/model/dit.py:8,270,274.
  * ``model/__init__.py`` imports ``tts.py`` which drags in the text encoder / HF transformers / Cython MAS.
    We only need ``model.diffusion`` so ``model`` is registered as a bare namespace package first.
"""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("DEX_REFERENCE_ROOT", "/root/reference")


def _install_timm_stub():
    if "timm" in sys.modules:
        return

    class Attention(nn.Module):
        def __init__(self, dim, num_heads=8, qkv_bias=False, **kw):
            super().__init__()
            self.num_heads = num_heads
            self.head_dim = dim // num_heads
            self.scale = self.head_dim ** -0.5
            self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
            self.proj = nn.Linear(dim, dim)

        def forward(self, x):
            B, N, C = x.shape
            qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
            q, k, v = qkv.unbind(0)
            attn = (q * self.scale) @ k.transpose(-2, -1)
            attn = attn.softmax(dim=-1)
            x = (attn @ v).transpose(1, 2).reshape(B, N, C)
            return self.proj(x)

    class Mlp(nn.Module):
        def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0, **kw):
            super().__init__()
            out_features = out_features or in_features
            hidden_features = hidden_features or in_features
            self.fc1 = nn.Linear(in_features, hidden_features)
            self.act = act_layer()
            self.fc2 = nn.Linear(hidden_features, out_features)

        def forward(self, x):
            return self.fc2(self.act(self.fc1(x)))

    class PatchEmbed(nn.Module):  # imported by name only, never instantiated on this path
        pass

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    vit = types.ModuleType("timm.models.vision_transformer")
    vit.Attention, vit.Mlp, vit.PatchEmbed = Attention, Mlp, PatchEmbed
    timm.models = models
    models.vision_transformer = vit
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.vision_transformer": vit})


class DotDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def load_reference(variant):
    """variant in {'dex','gedex'} -> the reference's ``model.diffusion`` module (fresh import)."""
    sub = {"dex": "DEX-TTS", "gedex": "GeDEX-TTS"}[variant]
    root = os.path.join(REF_ROOT, sub)
    _install_timm_stub()
    for name in [m for m in sys.modules if m == "model" or m.startswith("model.")]:
        del sys.modules[name]
    pkg = types.ModuleType("model")
    pkg.__path__ = [os.path.join(root, "model")]
    sys.modules["model"] = pkg
    return importlib.import_module("model.diffusion")


def build_reference_decoder(variant, decoder_cfg, dit_cfg, n_feats=80, n_spks=None, spk_emb_dim=64):
    """Construct the reference ``Diffusion`` exactly as tts.py does (DEX-TTS/model/tts.py:30,
    GeDEX-TTS/model/tts.py:25)."""
    mod = load_reference(variant)
    if n_spks is None:
        n_spks = 0 if variant == "dex" else 1
    dec = mod.Diffusion(**decoder_cfg, dit_cfg=DotDict(dit_cfg), n_feats=n_feats, n_spks=n_spks,
                        spk_emb_dim=spk_emb_dim)
    return dec.eval(), mod


def load_reference_text_encoder(variant="dex"):
    """-> the reference's ``model.text_encoder`` module (DEX-TTS or GeDEX-TTS), imported unmodified.  Extra shims (none edits a
    reference file): ``transformers.top_k_top_p_filtering`` (imported by name at DEX-TTS/model/retention.py:12, removed from
    transformers after the pinned 4.35.2, never called on this path) and ``timm.models.layers.drop_path`` (retention.py:9; the
    identity in eval mode).  transformers is imported before the timm stub is installed: its lazy-module probe calls
    ``importlib.util.find_spec('timm')``, which rejects a stub without a ``__spec__``."""
    import transformers
    if not hasattr(transformers, "top_k_top_p_filtering"):
        transformers.top_k_top_p_filtering = lambda *a, **k: None
    load_reference(variant)
    import timm
    layers = types.ModuleType("timm.models.layers")
    layers.drop_path = lambda x, drop_prob=0.0, training=False, scale_by_keep=True: x
    sys.modules["timm.models.layers"] = layers
    timm.models.layers = layers
    return importlib.import_module("model.text_encoder")


def build_reference_text_encoder(encoder_cfg, n_vocab=149, n_feats=80, n_spks=1, spk_emb_dim=64, variant="dex"):
    """Construct the reference ``TextEncoder`` as DEX-TTS/model/tts.py:29 (GeDEX-TTS/model/tts.py:24) does.  transformers 5.x ``PretrainedConfig`` no longer
    defines ``use_cache`` (read at DEX-TTS/model/retnet.py:79): set it on the instance's config, as SURVEY.md §8c prescribes."""
    mod = load_reference_text_encoder(variant)
    enc = mod.TextEncoder(**encoder_cfg, n_vocab=n_vocab, n_feats=n_feats, n_spks=n_spks, spk_emb_dim=spk_emb_dim)
    if not hasattr(enc.encoder.config, "use_cache"):
        enc.encoder.config.use_cache = True
    return enc.eval(), mod


def load_reference_tts(variant="dex"):
    """-> the reference's ``model.tts`` module (``DeXTTS`` / ``GeDEXTTS``), imported unmodified.  On top of the text-encoder shims:
    ``model.monotonic_align`` (DEX-TTS/model/tts.py:7; a Cython extension whose shipped binary is stale, called only by
    ``compute_loss`` at tts.py:108) is registered as a stub whose ``maximum_path`` raises."""
    load_reference_text_encoder(variant)
    mas = types.ModuleType("model.monotonic_align")

    def maximum_path(*a, **k):
        raise RuntimeError("monotonic_align is a training-only Cython extension; not available in the oracle harness")
    mas.maximum_path = maximum_path
    sys.modules["model.monotonic_align"] = mas
    sys.modules["model"].monotonic_align = mas
    return importlib.import_module("model.tts")


def reference_model_cfg(variant="dex", n_vocab=149, dataset=None):
    """``cfg.model`` as the entry scripts build it: the ``model:`` block of config/{VCTK,LJSpeech}/base.yaml with attribute access,
    plus ``n_vocab`` = len(symbols) + 1 (add_blank), which main.py / synthesize.py fill in at run time."""
    import yaml
    sub, ds = {"dex": ("DEX-TTS", "VCTK"), "gedex": ("GeDEX-TTS", "LJSpeech")}[variant]
    if dataset is not None:
        ds = dataset                                             # e.g. DEX-TTS/config/LibriTTS/base.yaml
    with open(os.path.join(REF_ROOT, sub, "config", ds, "base.yaml")) as f:
        raw = yaml.safe_load(f)["model"]
    wrap = lambda d: DotDict({k: wrap(v) if isinstance(v, dict) else v for k, v in d.items()})
    cfg = wrap(raw)
    cfg.n_vocab = n_vocab
    return cfg


def build_reference_tts(variant="dex", n_vocab=149, dataset=None):
    """Construct ``DeXTTS(cfg.model)`` / ``GeDEXTTS(cfg.model)`` as synthesize.py:67 does (eval mode)."""
    mod = load_reference_tts(variant)
    cfg = reference_model_cfg(variant, n_vocab, dataset)
    model = (mod.DeXTTS if variant == "dex" else mod.GeDEXTTS)(cfg)
    if not hasattr(model.encoder.encoder.config, "use_cache"):
        model.encoder.encoder.config.use_cache = True
    return model.eval(), mod, cfg
