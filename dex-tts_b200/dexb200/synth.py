"""Deterministic synthetic weights and inputs (no checkpoints or datasets exist offline).

``synth_decoder_weights`` draws every decoder tensor from a generator seeded by (seed, crc32(name)) so the same
numbers are produced wherever the manifest is evaluated (build container, GPU box, any rank) and independent of
module construction order.  Tensors the reference zero-initialises (adaLN-Zero, final DiT linear, Rezero gates,
freq pos-embed) are drawn NON-zero here: at the reference's default init the whole DiT / adaptor / linear-attention
branch multiplies by zero and a parity test would exercise ~40 % of the FLOPs only.
"""
import zlib

import torch

from .manifest import decoder_manifest


def _gen(seed, name):
    g = torch.Generator()
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
    return g


def synth_decoder_weights(cfg, seed=100, live=True, prefix="denoise_fn."):
    """-> {prefix+name: fp32 CPU tensor} for every manifest entry."""
    out = {}
    for e in decoder_manifest(cfg):
        g = _gen(seed, e.name)
        bound = 1.0 / max(e.fan_in, 1) ** 0.5
        if e.init in ("conv", "lin"):
            t = (torch.rand(e.shape, generator=g) * 2 - 1) * bound
        elif e.init == "posconv":
            if e.name.endswith(".bias"):
                t = torch.randn(e.shape, generator=g) * 0.02 if live else torch.zeros(e.shape)
            else:
                k, c = e.shape[2], cfg.hidden
                t = torch.randn(e.shape, generator=g) * (4.0 / (k * c)) ** 0.5
        elif e.init == "one":
            t = 1.0 + (0.1 * torch.randn(e.shape, generator=g) if live else torch.zeros(e.shape))
        elif e.init == "zero":
            if not live:
                t = torch.zeros(e.shape)
            elif e.name.endswith(".fn.g"):
                t = torch.full(e.shape, 0.5)
            elif e.name.endswith("block.1.bias"):
                t = 0.1 * torch.randn(e.shape, generator=g)
            elif "adaLN" in e.name or "final_layer.linear" in e.name:
                t = (torch.rand(e.shape, generator=g) * 2 - 1) * (bound if e.name.endswith("weight") else 0.1)
            else:
                t = 0.02 * torch.randn(e.shape, generator=g)
        else:
            raise ValueError(e.init)
        out[prefix + e.name] = t.float().contiguous()
    return out


def synth_inputs(cfg, B, T, Ts=259, seed=1234, ragged=False):
    """Seeded synthetic decoder inputs of SURVEY.md §8(d): mu ~ N(0,1) (B,80,T), noise z ~ N(0,1) drawn on CPU,
    lengths (all T, or ragged in [0.6T, T]); DEX conditioning sty (B,2*dim,Ts), 6 ref skips (B,2*dim,Ts)."""
    g = torch.Generator()
    g.manual_seed(seed)
    nf = cfg.n_feats
    mu = torch.randn(B, nf, T, generator=g)
    z = torch.randn(B, nf, T, generator=g)
    if ragged:
        lens = (torch.rand(B, generator=g) * 0.4 + 0.6) * T
        lens = lens.long().clamp(1, T)
        lens[0] = T
    else:
        lens = torch.full((B,), T, dtype=torch.long)
    mask = (torch.arange(T)[None, :] < lens[:, None]).float().unsqueeze(1)
    mu = mu * mask                                   # mu_y is exactly zero on padded frames upstream (attn path)
    out = dict(mu=mu, z=z, y_lengths=lens, mask=mask)
    if cfg.variant == "dex":
        c = 2 * cfg.dim
        sty = torch.randn(B, c, Ts, generator=g)
        refs = [torch.randn(B, c, Ts, generator=g) for _ in range(6)]
        if ragged:
            sl = ((torch.rand(B, generator=g) * 0.4 + 0.6) * Ts).long().clamp(1, Ts)
            sl[0] = Ts
        else:
            sl = torch.full((B,), Ts, dtype=torch.long)
        smask = (torch.arange(Ts)[None, :] < sl[:, None]).float().unsqueeze(1)
        out.update(sty=sty * smask, sty_lengths=sl, ref_skips=[r * smask for r in refs], ref_lengths=sl.clone())
    elif cfg.n_spks > 1:                           # GeDEX-TTS speaker embedding (rows of nn.Embedding(n_spks, spk_emb_dim))
        out["spk"] = torch.randn(B, cfg.spk_emb_dim, generator=g)
    return out


def tiv_manifest(c_in=80, c_out=64, num_layer=6, c_h=128):
    """[(name relative to ``tiv_encoder.``, shape, kind)] -- the ``state_dict`` of the reference TIVEncoder
    (DEX-TTS/model/ref_encoder.py:83-93 over BasicConv, DEX-TTS/model/base.py:33-50), in its own order."""
    out = []

    def basic(p, ci, co, bn):
        out.append((p + ".conv.weight", (co, ci, 3), "conv"))
        if bn:
            out.extend([(p + ".bn.weight", (co,), "bn_w"), (p + ".bn.bias", (co,), "bn_b"),
                        (p + ".bn.running_mean", (co,), "bn_rm"), (p + ".bn.running_var", (co,), "bn_rv"),
                        (p + ".bn.num_batches_tracked", (), "bn_n")])
    basic("in_conv", c_in, c_h, True)
    for i in range(num_layer):
        basic(f"conv_blocks.{i}.conv_block.0", c_h, c_h, True)
        basic(f"conv_blocks.{i}.conv_block.1", c_h, c_h, False)
    basic("out_conv", c_h, c_out, True)
    return out


def synth_tiv_weights(c_in=80, c_out=64, num_layer=6, c_h=128, seed=100, prefix="tiv_encoder."):
    """Seeded TIV-encoder tensors.  The BatchNorm running statistics are drawn away from their (0, 1) initial values so the
    eval-mode normalisation is really exercised."""
    out = {}
    for name, shape, kind in tiv_manifest(c_in, c_out, num_layer, c_h):
        g = _gen(seed, "tiv_encoder." + name)
        if kind == "conv":
            bound = 1.0 / (shape[1] * shape[2]) ** 0.5
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound * 1.7
        elif kind == "bn_w":
            t = 1.0 + 0.2 * torch.randn(shape, generator=g)
        elif kind == "bn_b":
            t = 0.2 * torch.randn(shape, generator=g)
        elif kind == "bn_rm":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_rv":
            t = 0.3 + 0.4 * torch.rand(shape, generator=g)
        else:
            t = torch.tensor(1000, dtype=torch.long)
        out[prefix + name] = t if kind == "bn_n" else t.float().contiguous()
    return out


def synth_ref_mel(B, T, n_feats=80, seed=77, ragged=False):
    """Seeded stand-in for a reference log-mel (B, n_feats, T) with lengths and mask (B,1,T)."""
    g = torch.Generator()
    g.manual_seed(seed)
    ref = torch.randn(B, n_feats, T, generator=g) * 1.5 - 4.0            # log-mel-like range
    if ragged:
        lens = ((torch.rand(B, generator=g) * 0.4 + 0.6) * T).long().clamp(2, T)
        lens[0] = T
    else:
        lens = torch.full((B,), T, dtype=torch.long)
    mask = (torch.arange(T)[None, :] < lens[:, None]).float().unsqueeze(1)
    return dict(ref=ref, ref_lengths=lens, mask=mask)


def tv_manifest(c_in=80, c_out=192, c_out_g=192, num_layer=6, c_h=128, n_emb=512):
    """[(name relative to ``tv_encoder.``, shape, kind)] -- the ``state_dict`` of the reference TVEncoder
    (DEX-TTS/model/ref_encoder.py:109-140 over BasicConv / Projection / VQEmbeddingEMA), in its own order."""
    out = []

    def basic(p, ci, co, norm):
        out.append((p + ".conv.weight", (co, ci, 3), "conv"))
        if norm == "ln":
            out.extend([(p + ".ln.weight", (co,), "bn_w"), (p + ".ln.bias", (co,), "bn_b")])
        elif norm == "bn":
            out.extend([(p + ".bn.weight", (co,), "bn_w"), (p + ".bn.bias", (co,), "bn_b"),
                        (p + ".bn.running_mean", (co,), "bn_rm"), (p + ".bn.running_var", (co,), "bn_rv"),
                        (p + ".bn.num_batches_tracked", (), "bn_n")])
    basic("in_conv", c_in, c_h, "ln")
    for i in range(num_layer):
        basic(f"conv_blocks.{i}.conv_block.0", c_h, c_h, "ln")
        basic(f"conv_blocks.{i}.conv_block.1", c_h, c_h, None)
    basic("out_conv", c_h, c_out, None)
    out.extend([("vq.embedding", (n_emb, c_out), "code"), ("vq.ema_count", (n_emb,), "ema_n"), ("vq.ema_weight", (n_emb, c_out), "code")])
    for j, (ci, k) in zip((1, 2), ((c_out, 3), (c_out_g, 3))):
        out.extend([(f"proj_0.conv_{j}.weight", (c_out_g, ci, k), "conv"), (f"proj_0.conv_{j}.bias", (c_out_g,), "bias")])
    # registration order of Projection.__init__ (ref_encoder.py:16-22): conv_1, norm_1, conv_2, norm_2, proj
    proj = [e for e in out if e[0].startswith("proj_0.")]
    out = [e for e in out if not e[0].startswith("proj_0.")]
    out.extend(proj[0:2])
    out.extend([("proj_0.norm_1.gamma", (c_out_g,), "bn_w"), ("proj_0.norm_1.beta", (c_out_g,), "bn_b")])
    out.extend(proj[2:4])
    out.extend([("proj_0.norm_2.gamma", (c_out_g,), "bn_w"), ("proj_0.norm_2.beta", (c_out_g,), "bn_b")])
    out.extend([("proj_0.proj.weight", (c_out_g, c_out_g, 1), "conv"), ("proj_0.proj.bias", (c_out_g,), "bias")])
    basic("proj_1", c_out_g, c_out_g, "bn")
    return out


def synth_tv_weights(c_in=80, c_out=192, c_out_g=192, num_layer=6, c_h=128, n_emb=512, seed=100, prefix="tv_encoder."):
    """Seeded TV-encoder tensors.  The codebook is drawn at the scale of the encoder output (the reference's 1/n_emb uniform
    init would put all 512 codes at the origin relative to the activations and make the argmin degenerate)."""
    out = {}
    for name, shape, kind in tv_manifest(c_in, c_out, c_out_g, num_layer, c_h, n_emb):
        g = _gen(seed, "tv_encoder." + name)
        if kind == "conv":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 1.7 / (shape[1] * shape[2]) ** 0.5
        elif kind == "bias":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_w":
            t = 1.0 + 0.2 * torch.randn(shape, generator=g)
        elif kind == "bn_b":
            t = 0.2 * torch.randn(shape, generator=g)
        elif kind == "bn_rm":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_rv":
            t = 0.3 + 0.4 * torch.rand(shape, generator=g)
        elif kind == "code":
            t = 0.6 * torch.randn(shape, generator=_gen(seed, "tv_encoder.vq.embedding"))     # ema_weight == embedding clone
        elif kind == "ema_n":
            t = torch.rand(shape, generator=g)
        else:
            t = torch.tensor(1000, dtype=torch.long)
        out[prefix + name] = t if kind == "bn_n" else t.float().contiguous()
    return out


def lf0_manifest(c_h=192, c_out=192, c_out_g=192, num_layer=2, c_in=1):
    """[(name relative to ``lf0_encoder.``, shape, kind)] -- the ``state_dict`` of the reference LF0Encoder
    (DEX-TTS/model/ref_encoder.py:36-56: BasicConv, nn.GRU(bidirectional), BasicConv, Projection), in its own order."""
    hid = c_h // 2
    out = [("in_conv.conv.weight", (c_h, c_in, 3), "conv"), ("in_conv.ln.weight", (c_h,), "bn_w"), ("in_conv.ln.bias", (c_h,), "bn_b")]
    for l in range(num_layer):
        for sfx in ("", "_reverse"):
            out.extend([(f"rnn_layer.weight_ih_l{l}{sfx}", (3 * hid, c_h), "gru"), (f"rnn_layer.weight_hh_l{l}{sfx}", (3 * hid, hid), "gru"),
                        (f"rnn_layer.bias_ih_l{l}{sfx}", (3 * hid,), "gru"), (f"rnn_layer.bias_hh_l{l}{sfx}", (3 * hid,), "gru")])
    out.extend([("out_conv.conv.weight", (c_out, c_h, 3), "conv"), ("out_conv.ln.weight", (c_out,), "bn_w"),
                ("out_conv.ln.bias", (c_out,), "bn_b")])
    out.extend([("proj.conv_1.weight", (c_out_g, c_out, 3), "conv"), ("proj.conv_1.bias", (c_out_g,), "bias"),
                ("proj.norm_1.gamma", (c_out_g,), "bn_w"), ("proj.norm_1.beta", (c_out_g,), "bn_b"),
                ("proj.conv_2.weight", (c_out_g, c_out_g, 3), "conv"), ("proj.conv_2.bias", (c_out_g,), "bias"),
                ("proj.norm_2.gamma", (c_out_g,), "bn_w"), ("proj.norm_2.beta", (c_out_g,), "bn_b"),
                ("proj.proj.weight", (c_out_g, c_out_g, 1), "conv"), ("proj.proj.bias", (c_out_g,), "bias")])
    return out


def synth_lf0_weights(c_h=192, c_out=192, c_out_g=192, num_layer=2, c_in=1, seed=100, prefix="lf0_encoder."):
    """Seeded LF0-encoder tensors (GRU tensors: nn.GRU's U(-1/sqrt(hidden), 1/sqrt(hidden)) init, slightly widened)."""
    out = {}
    hid = c_h // 2
    for name, shape, kind in lf0_manifest(c_h, c_out, c_out_g, num_layer, c_in):
        g = _gen(seed, "lf0_encoder." + name)
        if kind == "conv":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 1.7 / (shape[1] * shape[2]) ** 0.5
        elif kind == "gru":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 1.5 / hid ** 0.5
        elif kind == "bias":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_w":
            t = 1.0 + 0.2 * torch.randn(shape, generator=g)
        else:
            t = 0.2 * torch.randn(shape, generator=g)
        out[prefix + name] = t.float().contiguous()
    return out


def synth_lf0(B, T, seed=55, ragged=False):
    """Seeded stand-in for a normalised log-F0 contour (B, T) with unvoiced (zero) stretches, lengths and mask (B,1,T)
    (lf0 ~ N(0,1) * (rand > 0.3))."""
    g = torch.Generator()
    g.manual_seed(seed)
    lf0 = torch.randn(B, T, generator=g) * (torch.rand(B, T, generator=g) > 0.3).float()
    if ragged:
        lens = ((torch.rand(B, generator=g) * 0.4 + 0.6) * T).long().clamp(1, T)
        lens[0] = T
    else:
        lens = torch.full((B,), T, dtype=torch.long)
    mask = (torch.arange(T)[None, :] < lens[:, None]).float().unsqueeze(1)
    return dict(lf0=lf0, lf0_lengths=lens, mask=mask)


def synth_conv_sty_weights(c_in=192, c_out=128, seed=100):
    """Seeded tensors of DeXTTS.conv_sty = nn.Conv1d(tv_encoder.c_out_g, 2 * decoder.dim, 1) (DEX-TTS/model/tts.py:31)."""
    w = (torch.rand(c_out, c_in, 1, generator=_gen(seed, "conv_sty.weight")) * 2 - 1) * 1.7 / c_in ** 0.5
    return {"conv_sty.weight": w.contiguous(), "conv_sty.bias": 0.1 * torch.randn(c_out, generator=_gen(seed, "conv_sty.bias"))}


def synth_align_inputs(B, Tx, n_feats=80, seed=61, ragged=False, mean_dur=4.0):
    """Text-encoder outputs for the duration / alignment glue (DEX-TTS/model/tts.py:52-56): mu_x (B, n_feats, Tx), logw and x_mask
    (B, 1, Tx).  Durations exp(logw) are drawn as integer + U(0.15, 0.85), so ceil() never sits on a rounding boundary of exp()."""
    g = _gen(seed, "align")
    x_lengths = torch.full((B,), Tx, dtype=torch.long)
    if ragged and B > 1:
        x_lengths[1:] = torch.randint(max(1, int(0.5 * Tx)), Tx + 1, (B - 1,), generator=g)
    x_mask = (torch.arange(Tx)[None, :] < x_lengths[:, None]).float().unsqueeze(1)
    dur = torch.floor(torch.rand(B, 1, Tx, generator=g) * 2 * mean_dur) + 0.15 + 0.7 * torch.rand(B, 1, Tx, generator=g)
    logw = torch.log(dur) + (1 - x_mask) * torch.randn(B, 1, Tx, generator=g)      # padded tokens carry arbitrary logw
    mu_x = torch.randn(B, n_feats, Tx, generator=g) * x_mask                       # TextEncoder: mu = proj_m(x) * x_mask
    return dict(logw=logw, x_mask=x_mask, mu_x=mu_x, x_lengths=x_lengths)


def text_manifest(n_vocab=149, n_feats=80, n_channels=192, filter_channels=1024, filter_channels_dp=256, n_heads=2, n_layers=8,
                  kernel_size=3, adaln=True, spk_emb_dim=0):
    """[(name relative to ``encoder.``, shape, kind)] -- the ``state_dict`` of the reference TextEncoder for n_spks <= 1
    (DEX-TTS/model/text_encoder.py:97-142: Embedding, ConvReluNorm prenet, RetNetModel, proj_m, DurationPredictor), in its order.
    adaln=False: GeDEX-TTS's encoder (no AdaptiveLayerNorm in the RetNet layers).  spk_emb_dim > 0: the n_spks > 1 layout, where the
    speaker embedding is concatenated to the prenet output and everything behind the prenet is n_channels + spk_emb_dim wide
    (text_encoder.py:119-127,135-136) -- oracle / fixtures only, the CUDA text encoder takes spk_emb_dim = 0."""
    P, Fc, Fd = n_channels, filter_channels, filter_channels_dp                      # P: embedding / prenet width
    out = [("emb.weight", (n_vocab, P), "emb")]
    for i in range(3):                                                               # prenet: kernel 5, 3 layers (:116-117)
        out.extend([(f"prenet.conv_layers.{i}.weight", (P, P, 5), "conv"), (f"prenet.conv_layers.{i}.bias", (P,), "bias")])
    for i in range(3):
        out.extend([(f"prenet.norm_layers.{i}.gamma", (P,), "bn_w"), (f"prenet.norm_layers.{i}.beta", (P,), "bn_b")])
    out.extend([("prenet.proj.weight", (P, P, 1), "conv"), ("prenet.proj.bias", (P,), "bias")])
    C = P + spk_emb_dim                                                              # width of everything behind the prenet
    for l in range(n_layers):                                                        # RetNetDecoderLayer (retention.py:396-513)
        p = f"encoder.layers.{l}."
        out.extend([(p + f"retention.{n}_proj.weight", (C, C), "lin") for n in ("q", "k", "v", "g", "out")])
        out.append((p + "retention_layer_norm.weight", (C,), "bn_w"))
        out.extend([(p + "ffn.fc1.weight", (Fc, C), "lin"), (p + "ffn.fc2.weight", (C, Fc), "lin"), (p + "ffn.gate.weight", (Fc, C), "lin"),
                    (p + "final_layer_norm.weight", (C,), "bn_w")])
        for a in (("adaln_1", "adaln_2") if adaln else ()):
            out.extend([(p + a + ".W_scale.weight", (C, C), "ada"), (p + a + ".W_scale.bias", (C,), "bn_w"),
                        (p + a + ".W_bias.weight", (C, C), "ada"), (p + a + ".W_bias.bias", (C,), "bn_b")])
    out.extend([("encoder.layer_norm.weight", (C,), "bn_w"),
                ("encoder.retnet_rel_pos.angle", (C // n_heads,), "angle"), ("encoder.retnet_rel_pos.decay", (n_heads,), "decay"),
                ("proj_m.weight", (n_feats, C, 1), "conv"), ("proj_m.bias", (n_feats,), "bias"),
                ("proj_w.conv_1.weight", (Fd, C, kernel_size), "conv"), ("proj_w.conv_1.bias", (Fd,), "bias"),
                ("proj_w.norm_1.gamma", (Fd,), "bn_w"), ("proj_w.norm_1.beta", (Fd,), "bn_b"),
                ("proj_w.conv_2.weight", (Fd, Fd, kernel_size), "conv"), ("proj_w.conv_2.bias", (Fd,), "bias"),
                ("proj_w.norm_2.gamma", (Fd,), "bn_w"), ("proj_w.norm_2.beta", (Fd,), "bn_b"),
                ("proj_w.proj.weight", (1, Fd, 1), "conv"), ("proj_w.proj.bias", (1,), "bias")])
    return out


def synth_text_weights(seed=100, prefix="encoder.", **dims):
    """Seeded TextEncoder tensors.  The tensors the reference zero-initialises (prenet.proj, the AdaLN W_scale / W_bias weights:
    text_encoder.py:52-53, base.py:174-178) are re-drawn so the style conditioning and the prenet residual branch are live; the two
    RetNetRelPos buffers keep the reference's closed forms (retention.py:75-86)."""
    man = text_manifest(**dims)
    n_heads = [shape[0] for _, shape, kind in man if kind == "decay"][0]
    out = {}
    for name, shape, kind in man:
        g = _gen(seed, "encoder." + name)
        if kind == "emb":
            t = torch.randn(shape, generator=g) * shape[1] ** -0.5
        elif kind == "conv":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 1.7 / (shape[1] * shape[2]) ** 0.5
        elif kind == "lin":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 1.7 / shape[1] ** 0.5
        elif kind == "ada":
            t = 0.02 * torch.randn(shape, generator=g)
        elif kind == "bias":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_w":
            t = 1.0 + 0.2 * torch.randn(shape, generator=g)
        elif kind == "bn_b":
            t = 0.2 * torch.randn(shape, generator=g)
        elif kind == "angle":
            a = 1.0 / (10000 ** torch.linspace(0, 1, shape[0] // 2))
            t = a.unsqueeze(-1).repeat(1, 2).flatten()
        else:
            t = torch.log(1 - 2 ** (-5 - torch.arange(n_heads, dtype=torch.float)))
        out[prefix + name] = t.float().contiguous()
    return out


def synth_text(B, Tx, n_vocab=149, c_sty=192, seed=81, ragged=False):
    """Seeded phoneme ids (B, Tx) in [0, n_vocab), lengths, and the style vector (B, c_sty) the AdaLN layers take
    (randint(0, 149) of length 128 / 512)."""
    g = _gen(seed, "text")
    x = torch.randint(0, n_vocab, (B, Tx), generator=g)
    x_lengths = torch.full((B,), Tx, dtype=torch.long)
    if ragged and B > 1:
        x_lengths[1:] = torch.randint(max(1, int(0.5 * Tx)), Tx + 1, (B - 1,), generator=g)
    x = x * (torch.arange(Tx)[None, :] < x_lengths[:, None])                       # the collate pads ids with 0
    return dict(x=x, x_lengths=x_lengths, sty=torch.randn(B, c_sty, generator=g))


LIBRITTS_MODEL_CFG = dict(                    # the `model:` block of DEX-TTS/config/LibriTTS/base.yaml:23-83 (+ n_vocab, set by the scripts)
    sil_token=True, add_blank=True, n_feats=80, n_spks=0, spk_emb_dim=64, n_vocab=149,
    tv_encoder=dict(c_in=80, num_layer=6, c_h=256, c_out=256, c_out_g=256, commit_w=0.25, n_emb=512),
    lf0_encoder=dict(c_in=1, c_h=256, c_out=256, c_out_g=256, num_layer=2),
    tiv_encoder=dict(c_in=80, num_layer=6, c_h=256, c_out=64),
    encoder=dict(n_channels=256, filter_channels=1024, filter_channels_dp=256, n_layers=8, kernel_size=3, p_dropout=0.1, n_heads=2,
                 window_size=4, use_softmax=True, use_decay=False),
    decoder=dict(dim=128, pe_scale=1000, dim_mults=[1, 2], model_type="dit", precond="edm", loss_type="base"),
    dit=dict(in_channels=3, patch_size=3, stride_size=2, overlap=True, hidden_size=384, depth=4, num_heads=2, mlp_ratio=2, out_channels=1,
             conv_pos=16, conv_pos_groups=8, use_decoder=False, mask_type="time_random"))


def synth_tts_weights(variant="dex", seed=100, dataset="VCTK"):
    """Every tensor of the reference ``DeXTTS`` / ``GeDEXTTS`` (n_spks <= 1) as one flat dict: the encoders and ``conv_sty`` under
    their ``state_dict`` names, the decoder under ``denoise_fn.*`` (upstream: ``decoder.denoise_fn.*`` and, aliased,
    ``decoder.precond_model.model.*``).  dataset="LibriTTS": the sizes of DEX-TTS/config/LibriTTS/base.yaml (decoder dim 128 / DiT
    hidden 384, 256-wide encoders)."""
    from .manifest import DecoderCfg
    if dataset == "LibriTTS":
        assert variant == "dex"
        m = LIBRITTS_MODEL_CFG
        w = dict(synth_decoder_weights(DecoderCfg.make("dex", dim=m["decoder"]["dim"], hidden=m["dit"]["hidden_size"]), seed=seed, live=True))
        e = m["encoder"]
        w.update(synth_text_weights(seed=seed, adaln=True, n_channels=e["n_channels"], filter_channels=e["filter_channels"],
                                    filter_channels_dp=e["filter_channels_dp"], n_heads=e["n_heads"], n_layers=e["n_layers"],
                                    kernel_size=e["kernel_size"]))
        tv, lf, ti = m["tv_encoder"], m["lf0_encoder"], m["tiv_encoder"]
        w.update(synth_tv_weights(c_in=tv["c_in"], c_out=tv["c_out"], c_out_g=tv["c_out_g"], num_layer=tv["num_layer"], c_h=tv["c_h"],
                                  n_emb=tv["n_emb"], seed=seed))
        w.update(synth_lf0_weights(c_h=lf["c_h"], c_out=lf["c_out"], c_out_g=lf["c_out_g"], num_layer=lf["num_layer"], c_in=lf["c_in"],
                                   seed=seed))
        w.update(synth_tiv_weights(c_in=ti["c_in"], c_out=ti["c_out"], num_layer=ti["num_layer"], c_h=ti["c_h"], seed=seed))
        w.update(synth_conv_sty_weights(c_in=tv["c_out_g"], c_out=2 * m["decoder"]["dim"], seed=seed))
        return w
    w = dict(synth_decoder_weights(DecoderCfg.make(variant), seed=seed, live=True))
    w.update(synth_text_weights(seed=seed, adaln=variant == "dex"))
    if variant == "dex":
        w.update(synth_tv_weights(seed=seed))
        w.update(synth_lf0_weights(seed=seed))
        w.update(synth_tiv_weights(seed=seed))
        w.update(synth_conv_sty_weights(seed=seed))
    return w


def reference_state_dict(w):
    """Flat dict of ``synth_tts_weights`` -> the key layout of the reference model's ``state_dict`` (decoder tensors twice)."""
    sd = {}
    for k, v in w.items():
        if k.startswith("denoise_fn."):
            sd["decoder." + k] = v
            sd["decoder.precond_model.model." + k[len("denoise_fn."):]] = v
        else:
            sd[k] = v
    return sd


def seeded_noise(seed):
    """noise(shape) -> N(0, 1) CPU tensor from a private generator: the injected Gaussian draw of Diffusion.forward."""
    g = torch.Generator()
    g.manual_seed(seed)
    randn = torch.randn                     # bound now: the golden generators patch torch.randn while the reference runs
    return lambda shape: randn(shape, generator=g)


HIFIGAN_V1_CFG = dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
                      resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]])    # DEX-TTS/hifigan/config.json


def synth_vocoder_weights(seed=100, cfg=None):
    """Seeded HiFi-GAN v1 generator tensors (state after remove_weight_norm) at a scale that keeps the activations O(1) through the four
    stages -- benchmark input for ``dexb200.hifigan`` (the parity fixtures use oracle/vocoder_oracle.py's own draw)."""
    from .hifigan.models import vocoder_manifest
    w = {}
    for name, shape in vocoder_manifest(cfg or HIFIGAN_V1_CFG):
        g = _gen(seed, "vocoder." + name)
        if name.endswith("bias"):
            w[name] = 0.05 * torch.randn(shape, generator=g)
        else:
            fan = (shape[0] * shape[2] / 8.0) if name.startswith("ups.") else shape[1] * shape[2]
            w[name] = torch.randn(shape, generator=g) * (0.7 / fan ** 0.5)
    return w
