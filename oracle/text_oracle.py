"""CPU oracle for the text side of DeXTTS.forward: TextEncoder (phoneme ids + style vector -> mu_x, logw, x_mask).

TEST INFRASTRUCTURE ONLY -- never imported by the product path (``dex-tts_b200/``); see the header of ``dex_oracle.py``.
SURVEY.md §8f rank 2.  The CUDA side of this stage is NOT built yet: this file and tests/golden/text_*.npz are the pinned
oracle the next round builds against (oracle first, then the boundary, then the kernels).

Functional restatement (plain torch CPU ops over a flat ``{name: tensor}`` dict keyed by the reference's ``state_dict`` names) of

    TextEncoder.forward                DEX-TTS/model/text_encoder.py:129-142
    ConvReluNorm / LayerNorm           DEX-TTS/model/text_encoder.py:11-64      (prenet: kernel 5, 3 layers, :116-117)
    DurationPredictor                  DEX-TTS/model/text_encoder.py:67-95
    RetNetModel.forward ('parallel')   DEX-TTS/model/retnet.py:57-176
    RetNetRelPos.forward ('parallel', use_decay=False)          DEX-TTS/model/retention.py:139-165
    RetNetDecoderLayer.forward         DEX-TTS/model/retention.py:458-514
    MultiScaleRetention (use_softmax)  DEX-TTS/model/retention.py:223-295       (rotary theta_shift :28-37, RMSNorm :50-69)
    GLU                                DEX-TTS/model/retention.py:346-381
    AdaptiveLayerNorm                  DEX-TTS/model/base.py:161-194

RetNetConfig values the reference leaves at their defaults (DEX-TTS/model/retnet_cfg.py:40-67): activation gelu (exact),
use_glu, subln -> pre-norm with alpha = 1, layernorm_eps 1e-6, no final embedding scale; eval mode, so every dropout / drop_path is
the identity.  With ``use_softmax=True, use_decay=False`` (DEX-TTS/config/VCTK/base.yaml:60-61) the "retention" is softmax
attention with rotary q / k, a key-padding fill of -1e4 (not -inf), a per-head RMS norm and a swish gate.

GeDEX-TTS ships the same encoder without the AdaptiveLayerNorm layers and without the ``sty`` argument (GeDEX-TTS/model/
text_encoder.py, retention.py): ``sty=None`` here.

Parity pin: outputs of the unmodified reference TextEncoder of both repositories, generated in the build container by oracle/make_golden_text.py and
committed as tests/golden/text_*.npz; tests/test_text_oracle.py replays them.
"""
import math

import torch
import torch.nn.functional as F


def _channel_layer_norm(x, gamma, beta, eps=1e-4):
    """text_encoder.py:21-30: normalise over the channel axis of (B, C, T), biased variance."""
    mean = x.mean(1, keepdim=True)
    var = ((x - mean) ** 2).mean(1, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * gamma.view(1, -1, 1) + beta.view(1, -1, 1)


def _rms_norm(x, weight=None, eps=1e-6):
    """retention.py:62-69."""
    y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    return y if weight is None else y * weight


def _rotary(x, sin, cos):
    """theta_shift, retention.py:28-37: pairs (x0, x1) -> (x0 cos - x1 sin, x1 cos + x0 sin), angle repeated per pair."""
    x1, x2 = x[..., ::2], x[..., 1::2]
    rot = torch.stack((-x2, x1), dim=-1).flatten(-2)
    return x * cos + rot * sin


def _ada_layer_norm(w, p, x, sty, eps=1e-5):
    """base.py:180-194: (x - mean) / sqrt(var + eps) over channels, then a style-predicted scale and bias."""
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    y = (x - mean) / (var + eps).sqrt()
    scale = F.linear(sty, w[p + ".W_scale.weight"], w[p + ".W_scale.bias"])
    bias = F.linear(sty, w[p + ".W_bias.weight"], w[p + ".W_bias.bias"])
    return y * scale.unsqueeze(1) + bias.unsqueeze(1)


def _retention(w, p, h, sin, cos, pair_mask, n_heads):
    """MultiScaleRetention.forward + parallel_retention with use_softmax (retention.py:223-295)."""
    B, T, C = h.shape
    d = C // n_heads
    heads = lambda t: t.view(B, T, n_heads, d).transpose(1, 2)
    q = heads(F.linear(h, w[p + ".q_proj.weight"]))
    k = heads(F.linear(h, w[p + ".k_proj.weight"])) * d ** -0.5                                       # :281
    v = heads(F.linear(h, w[p + ".v_proj.weight"]))
    g = F.linear(h, w[p + ".g_proj.weight"])
    s = (_rotary(q, sin, cos) @ _rotary(k, sin, cos).transpose(-1, -2)) * pair_mask                   # :238-239
    s = s.masked_fill(pair_mask == 0, -1e4)                                                           # :242
    o = (F.softmax(s, dim=-1) @ v).transpose(1, 2)                                                    # :243-249  (B, T, H, d)
    o = _rms_norm(o).reshape(B, T, C)                                                                 # group_norm :290
    return F.linear(F.silu(g) * o, w[p + ".out_proj.weight"])                                         # :292-293


def _glu(w, p, x):
    """GLU.forward, retention.py:371-381: fc2(gelu(fc1 x) * gate x), no biases."""
    return F.linear(F.gelu(F.linear(x, w[p + ".fc1.weight"])) * F.linear(x, w[p + ".gate.weight"]), w[p + ".fc2.weight"])


def retnet(w, h, x_mask, sty, n_layers=8, n_heads=2, prefix="encoder.encoder", taps=None):
    """RetNetModel.forward(inputs_embeds=h (B,T,C), attention_mask=x_mask (B,1,T), sty (B,C)) -> last_hidden_state (B,T,C).
    sty=None: the GeDEX-TTS flavour, whose decoder layers have no AdaptiveLayerNorm (GeDEX-TTS/model/retention.py:453-505)."""
    T = h.shape[1]
    index = torch.arange(T).to(h)
    angle = w[prefix + ".retnet_rel_pos.angle"]
    sin, cos = torch.sin(index[:, None] * angle[None, :]), torch.cos(index[:, None] * angle[None, :])  # retention.py:140-142
    pair_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)                                            # :143   (B, 1, T, T)
    for l in range(n_layers):
        p = f"{prefix}.layers.{l}"
        h = h + _retention(w, p + ".retention", _rms_norm(h, w[p + ".retention_layer_norm.weight"]), sin, cos, pair_mask, n_heads)
        if sty is not None:
            h = _ada_layer_norm(w, p + ".adaln_1", h, sty)                                            # :487 (DEX-TTS only)
        h = h + _glu(w, p + ".ffn", _rms_norm(h, w[p + ".final_layer_norm.weight"]))
        if sty is not None:
            h = _ada_layer_norm(w, p + ".adaln_2", h, sty)                                            # :505 (DEX-TTS only)
        if taps is not None:
            taps[f"layer{l}"] = h
    return _rms_norm(h, w[prefix + ".layer_norm.weight"])                                             # retnet.py:162-163


def prenet(w, x, x_mask, prefix="encoder.prenet", n_layers=3):
    """ConvReluNorm.forward, text_encoder.py:56-64."""
    x_org = x
    for i in range(n_layers):
        x = F.conv1d(x * x_mask, w[f"{prefix}.conv_layers.{i}.weight"], w[f"{prefix}.conv_layers.{i}.bias"], padding=2)
        x = F.relu(_channel_layer_norm(x, w[f"{prefix}.norm_layers.{i}.gamma"], w[f"{prefix}.norm_layers.{i}.beta"]))
    return (x_org + F.conv1d(x, w[prefix + ".proj.weight"], w[prefix + ".proj.bias"])) * x_mask


def duration_predictor(w, x, x_mask, prefix="encoder.proj_w"):
    """DurationPredictor.forward, text_encoder.py:84-95 (ReLU before the norm here, unlike the prenet)."""
    for i in (1, 2):
        cw = w[f"{prefix}.conv_{i}.weight"]
        x = F.relu(F.conv1d(x * x_mask, cw, w[f"{prefix}.conv_{i}.bias"], padding=cw.shape[-1] // 2))
        x = _channel_layer_norm(x, w[f"{prefix}.norm_{i}.gamma"], w[f"{prefix}.norm_{i}.beta"])
    return F.conv1d(x * x_mask, w[prefix + ".proj.weight"], w[prefix + ".proj.bias"]) * x_mask


def text_encoder(w, x_ids, x_lengths, sty, n_layers=8, n_heads=2, prefix="encoder", taps=None, spk=None):
    """TextEncoder.forward(x, x_lengths, sty, spk=None) for n_spks <= 1, text_encoder.py:129-142; with sty=None it is
    GeDEX-TTS's TextEncoder.forward(x, x_lengths, spk=None) (GeDEX-TTS/model/text_encoder.py:132-146: the same code without the style).
    spk (B, spk_emb_dim): the n_spks > 1 branch -- the speaker embedding is repeated over time and concatenated to the prenet output
    (:135-136), everything behind it is C + spk_emb_dim wide.
    x_ids (B,Tx) long, x_lengths (B,), sty (B,C) -> (mu_x (B,n_feats,Tx), logw (B,1,Tx), x_mask (B,1,Tx))."""
    emb = w[prefix + ".emb.weight"]
    C = emb.shape[1]
    x = (emb[x_ids] * math.sqrt(C)).transpose(1, -1)                                                  # :130-131
    x_mask = (torch.arange(x.shape[2])[None, :] < x_lengths[:, None]).to(x.dtype).unsqueeze(1)        # :132
    x = prenet(w, x, x_mask, prefix + ".prenet")                                                      # :134
    if taps is not None:
        taps["prenet"] = x
    if spk is not None:
        x = torch.cat([x, spk.unsqueeze(-1).repeat(1, 1, x.shape[-1])], dim=1)                        # :135-136
    x = retnet(w, x.transpose(1, 2), x_mask, sty, n_layers, n_heads, prefix + ".encoder", taps).transpose(1, 2) * x_mask   # :137
    mu = F.conv1d(x, w[prefix + ".proj_m.weight"], w[prefix + ".proj_m.bias"]) * x_mask               # :138
    return mu, duration_predictor(w, x, x_mask, prefix + ".proj_w"), x_mask                           # :140-142
