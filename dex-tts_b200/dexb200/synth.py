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
