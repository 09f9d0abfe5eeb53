"""Parameter manifest of the reverse-diffusion decoder.

One table drives everything that must agree on parameter names and shapes:
  * ``model.diffusion.Diffusion`` builds its ``nn.Parameter`` tree from it, so ``state_dict()`` keys equal the
    reference's (DEX-TTS/model/diffusion.py:122-175,238-243, DEX-TTS/model/dit.py:328-396,
    DEX-TTS/model/ref_encoder.py:142-152,239-262) and upstream checkpoints load with ``strict=True``;
  * ``dexb200.engine`` packs exactly these tensors for the CUDA path;
  * ``dexb200.synth`` draws deterministic, fully "live" test/bench weights for them.

Each entry: (name relative to ``denoise_fn``, shape, init) with init in
  'conv'/'lin'  U(-1/sqrt(fan_in), 1/sqrt(fan_in))  (torch default for Conv2d / Linear weight and bias)
  'one' / 'zero'  constants (GroupNorm affine; adaLN-Zero, final DiT linear, Rezero gate: DEX-TTS/model/dit.py:398-407,
                  DEX-TTS/model/diffusion.py:38)
  'posconv'       N(0, sqrt(4/(k*e)))  (DEX-TTS/model/dit.py:84-85)
"""
from collections import namedtuple

Entry = namedtuple("Entry", "name shape init fan_in")


class DecoderCfg(dict):
    """Hyper-parameters the decoder is built from (yaml ``model.decoder`` + ``model.dit``)."""
    __getattr__ = dict.__getitem__

    @staticmethod
    def make(variant="dex", dim=64, hidden=256, depth=4, heads=2, mlp_ratio=2, patch=None, stride=None,
             conv_pos=16, conv_pos_groups=8, n_feats=80, pe_scale=1000, n_spks=None, spk_emb_dim=64):
        assert variant in ("dex", "gedex")
        if patch is None:
            patch = 3 if variant == "dex" else 7
        if stride is None:
            stride = 2 if variant == "dex" else 4
        if n_spks is None:
            n_spks = 0 if variant == "dex" else 1
        return DecoderCfg(variant=variant, dim=dim, hidden=hidden, depth=depth, heads=heads, mlp_ratio=mlp_ratio,
                          patch=patch, stride=stride, conv_pos=conv_pos, conv_pos_groups=conv_pos_groups,
                          n_feats=n_feats, pe_scale=pe_scale, n_spks=n_spks, spk_emb_dim=spk_emb_dim)


def _conv(out, name, co, ci, kh, kw, bias=True, groups=1, init="conv"):
    fan = (ci // groups) * kh * kw
    out.append(Entry(name + ".weight", (co, ci // groups, kh, kw), init, fan))
    if bias:
        out.append(Entry(name + ".bias", (co,), "zero" if init in ("zero", "posconv") else "conv", fan))


def _lin(out, name, co, ci, bias=True, init="lin"):
    out.append(Entry(name + ".weight", (co, ci), init, ci))
    if bias:
        out.append(Entry(name + ".bias", (co,), init, ci))


def _block(out, name, ci, co):
    _conv(out, name + ".block.0", co, ci, 3, 3)
    out.append(Entry(name + ".block.1.weight", (co,), "one", 1))
    out.append(Entry(name + ".block.1.bias", (co,), "zero", 1))


def _resnet(out, name, ci, co, tdim):
    _lin(out, name + ".mlp.1", co, tdim)
    _block(out, name + ".block1", ci, co)
    _block(out, name + ".block2", co, co)
    if ci != co:
        _conv(out, name + ".res_conv", co, ci, 1, 1)


def _linattn(out, name, c, heads=4, dim_head=32):
    out.append(Entry(name + ".fn.g", (1,), "zero", 1))
    _conv(out, name + ".fn.fn.to_qkv", 3 * heads * dim_head, c, 1, 1, bias=False)
    _conv(out, name + ".fn.fn.to_out", c, heads * dim_head, 1, 1)


def decoder_manifest(cfg):
    """Ordered list of Entry for ``Diffusion.denoise_fn`` (order = the reference's registration order)."""
    d, hid = cfg.dim, cfg.hidden
    dex = cfg.variant == "dex"
    cin = 2 + (1 if cfg.n_spks > 1 else 0)
    mid = 2 * d
    out = []
    _lin(out, "mlp.0", 4 * d, d)
    _lin(out, "mlp.2", d, 4 * d)
    if dex:
        for m in ("mlp_adap", "mlp_adap_sty"):
            _lin(out, m + ".0", d, d)
            _lin(out, m + ".2", 2 * d, d)
    if cfg.n_spks > 1:
        _lin(out, "spk_mlp.0", 4 * cfg.spk_emb_dim, cfg.spk_emb_dim)
        _lin(out, "spk_mlp.2", cfg.n_feats, 4 * cfg.spk_emb_dim)
    _resnet(out, "downs.0.0", cin, d, d)
    _resnet(out, "downs.0.1", d, d, d)
    _linattn(out, "downs.0.2", d)
    _conv(out, "downs.0.3.conv", d, d, 3, 3)
    _resnet(out, "downs.1.0", d, mid, d)
    _resnet(out, "downs.1.1", mid, mid, d)
    _linattn(out, "downs.1.2", mid)
    _resnet(out, "ups.0.0", 2 * mid, d, d)
    _resnet(out, "ups.0.1", d, d, d)
    _linattn(out, "ups.0.2", d)
    out.append(Entry("ups.0.3.conv.weight", (d, d, 4, 4), "conv", d * 16))      # ConvTranspose2d: (in, out, kh, kw)
    out.append(Entry("ups.0.3.conv.bias", (d,), "conv", d * 16))
    if dex:
        for nm in ("w_q", "w_k", "w_v", "linear"):
            _lin(out, "tv_adaptor." + nm, mid, mid, bias=False)
        _lin(out, "tiv_adaptor.mean_sap.W", 1, mid)
        _lin(out, "tiv_adaptor.std_sap.W", 1, mid)
    fq = (cfg.n_feats // 2) // cfg.stride                                     # PatchEmbed2D.grid_size[0], dit.py:46
    out.append(Entry("vit.freq_new_pos_embed", (1, hid, fq, 1), "zero", 1))
    _conv(out, "vit.x_embedder.proj.0", mid, mid, cfg.patch, cfg.patch, groups=mid)
    _conv(out, "vit.x_embedder.proj.2", hid, mid, 1, 1)
    _lin(out, "vit.t_embedder.mlp.0", hid, 256)
    _lin(out, "vit.t_embedder.mlp.2", hid, hid)
    _conv(out, "vit.pos_conv.0", hid, hid, cfg.conv_pos, cfg.conv_pos, groups=cfg.conv_pos_groups, init="posconv")
    mh = int(hid * cfg.mlp_ratio)
    for i in range(cfg.depth):
        b = f"vit.blocks.{i}"
        _lin(out, b + ".attn.qkv", 3 * hid, hid)
        _lin(out, b + ".attn.proj", hid, hid)
        _lin(out, b + ".mlp.fc1", mh, hid)
        _lin(out, b + ".mlp.fc2", hid, mh)
        _lin(out, b + ".adaLN_modulation.1", 6 * hid, hid, init="zero")
    _lin(out, "vit.final_layer.linear", cfg.stride * cfg.stride * mid, hid, init="zero")
    _lin(out, "vit.final_layer.adaLN_modulation.1", 2 * hid, hid, init="zero")
    _block(out, "final_block", d, d)
    _conv(out, "final_conv", 1, d, 1, 1)
    return out
