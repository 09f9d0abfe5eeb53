"""CPU oracle for the DEX-TTS / GeDEX-TTS reverse-diffusion hot path.

TEST INFRASTRUCTURE ONLY -- never imported by the product path (``dex-tts_b200/``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it,
and there only as the checker / the timed CPU baseline.

It is a functional restatement (plain torch CPU ops over a flat ``{name: tensor}`` weight dict keyed by the
reference's own ``state_dict`` names) of:

    Diffusion.forward(infer=True)      DEX-TTS/model/diffusion.py:250-259   (GeDEX-TTS/model/diffusion.py:220-229)
    ablation_sampler (euler/edm/linear) DEX-TTS/model/edm.py:104-211         (GeDEX-TTS/model/edm.py:109-216)
    EDMPrecond.forward                 DEX-TTS/model/edm.py:88-98
    DiffusionDenoiser.forward          DEX-TTS/model/diffusion.py:190-236   (GeDEX-TTS/model/diffusion.py:168-207)
    DiTMask.forward                    DEX-TTS/model/dit.py:479-519
    TVAdaptor / TIVAdaptor             DEX-TTS/model/ref_encoder.py:142-179, 239-273
    TIVEncoder.forward (eval)          DEX-TTS/model/ref_encoder.py:83-107  (pre-loop stage, SURVEY.md §8f rank 1;
                                       pinned by tests/golden/tiv_*.npz from oracle/make_golden_tiv.py)
    TVEncoder.forward (eval)           DEX-TTS/model/ref_encoder.py:109-140 (same stage; tests/golden/tv_*.npz)
    LF0Encoder.forward (eval)          DEX-TTS/model/ref_encoder.py:36-56   (same stage; tests/golden/lf0_*.npz)
    style fusion + conv_sty            DEX-TTS/model/tts.py:45-49           (same fixtures)
    duration / alignment glue          DEX-TTS/model/tts.py:55-68, DEX-TTS/model/utils.py:6-39  (first
                                       piece; tests/golden/align_*.npz from oracle/make_golden_align.py)

Third-party arithmetic on the path: ``timm`` (un-pinned, DEX-TTS/requirements.txt:16) ``Attention`` and ``Mlp``
(call sites DEX-TTS/model/dit.py:270,274); their published algorithm is restated in ``_dit_block``.

Parity pin: the reference ships no tests or golden vectors, so this oracle is pinned against
outputs of the reference itself, generated in the build container by ``oracle/make_golden.py`` (which imports
the unmodified reference from /root/reference) and committed under ``tests/golden/``;
``tests/test_oracle_golden.py`` replays them.

Batched DEX note: the reference sampler hands a 0-dim sigma to the net, which crashes for B>1 in the
TV adaptor.  The oracle broadcasts sigma over the batch, which is bit-identical to the
reference at B=1 and is what ``make_golden.py`` does to the reference (sigma.expand(B)) for B>1.
"""
import math

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# configuration
# ----------------------------------------------------------------------------------------------

def make_cfg(variant="dex", dim=64, hidden=256, depth=4, heads=2, mlp_ratio=2, patch=None, stride=None,
             conv_pos=16, conv_pos_groups=8, n_feats=80, pe_scale=1000, n_spks=None):
    """Hyper-parameters of the decoder (DEX-TTS/config/VCTK/base.yaml:56-78, GeDEX-TTS/config/LJSpeech/base.yaml:41-63).
    n_spks > 1 (GeDEX-TTS only) adds the speaker channel of GeDEX-TTS/model/diffusion.py:132-134,170-175."""
    assert variant in ("dex", "gedex")
    if patch is None:
        patch = 3 if variant == "dex" else 7
    if stride is None:
        stride = 2 if variant == "dex" else 4
    if n_spks is None:
        n_spks = 0 if variant == "dex" else 1
    return dict(variant=variant, dim=dim, hidden=hidden, depth=depth, heads=heads, mlp_ratio=mlp_ratio,
                patch=patch, stride=stride, conv_pos=conv_pos, conv_pos_groups=conv_pos_groups,
                n_feats=n_feats, pe_scale=pe_scale, n_spks=n_spks)


def sequence_mask(lengths, max_len):
    """DEX-TTS/model/utils.py:6-10."""
    ar = torch.arange(int(max_len), dtype=lengths.dtype, device=lengths.device)
    return ar.unsqueeze(0) < lengths.unsqueeze(1)


# ----------------------------------------------------------------------------------------------
# elementary pieces
# ----------------------------------------------------------------------------------------------

def mish(x):
    """DEX-TTS/model/diffusion.py:11-13."""
    return x * torch.tanh(F.softplus(x))


def sinusoidal_emb(t, dim, scale):
    """SinusoidalPosEmb, DEX-TTS/model/diffusion.py:113-120 (sin first, then cos)."""
    half = dim // 2
    k = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half, dtype=torch.float32) * -k).to(t.dtype)
    arg = scale * t.unsqueeze(1) * freqs.unsqueeze(0)
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


def timestep_embedding(t, dim, max_period=10000):
    """TimestepEmbedder.timestep_embedding, DEX-TTS/model/dit.py:233-251 (cos first, then sin; raw t, no scale)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.dtype)
    arg = t[:, None] * freqs[None]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


def _lin(w, p, x, bias=True):
    return F.linear(x, w[p + ".weight"], w[p + ".bias"] if bias else None)


def _block(w, p, x, mask, groups=8):
    """Block: conv3x3 -> GroupNorm(8) -> Mish, masked in and out.  DEX-TTS/model/diffusion.py:44-53."""
    y = F.conv2d(x * mask, w[p + ".block.0.weight"], w[p + ".block.0.bias"], padding=1)
    y = F.group_norm(y, groups, w[p + ".block.1.weight"], w[p + ".block.1.bias"], eps=1e-5)
    return mish(y) * mask


def _resnet(w, p, x, mask, temb):
    """ResnetBlock, DEX-TTS/model/diffusion.py:56-74."""
    h = _block(w, p + ".block1", x, mask)
    h = h + _lin(w, p + ".mlp.1", mish(temb))[:, :, None, None]
    h = _block(w, p + ".block2", h, mask)
    if (p + ".res_conv.weight") in w:
        r = F.conv2d(x * mask, w[p + ".res_conv.weight"], w[p + ".res_conv.bias"])
    else:
        r = x * mask
    return h + r


def _linear_attention(w, p, x, heads=4, dim_head=32):
    """Residual(Rezero(LinearAttention)), DEX-TTS/model/diffusion.py:34-41,77-105.
    softmax of k runs over ALL h*w positions, padding included (no mask)."""
    b, c, hh, ww = x.shape
    qkv = F.conv2d(x, w[p + ".fn.fn.to_qkv.weight"])
    qkv = qkv.reshape(b, 3, heads, dim_head, hh * ww)          # 'b (qkv heads c) h w -> qkv b heads c (h w)'
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    k = k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q)
    out = out.reshape(b, heads * dim_head, hh, ww)
    out = F.conv2d(out, w[p + ".fn.fn.to_out.weight"], w[p + ".fn.fn.to_out.bias"])
    return out * w[p + ".fn.g"] + x


def _instance_norm_2d(x, eps=1e-5):
    """InstanceNorm2D, DEX-TTS/model/base.py:90-114: UNBIASED variance over all h*w, padding included."""
    n, c = x.shape[:2]
    flat = x.reshape(n, c, -1)
    std = (flat.var(dim=2) + eps).sqrt().reshape(n, c, 1, 1)
    mean = flat.mean(dim=2).reshape(n, c, 1, 1)
    return (x - mean) / std


def ref_stats(ref_skips, eps=1e-5):
    """DiffusionDenoiser._stack_stats + InstanceNorm1D.cal_stats (lengths ignored!),
    DEX-TTS/model/diffusion.py:177-188, DEX-TTS/model/base.py:72-78.  -> (B,L,C), (B,L,C)."""
    means = [r.mean(-1) for r in ref_skips]
    stds = [(r.var(-1) + eps).sqrt() for r in ref_skips]
    return torch.stack(means, dim=1), torch.stack(stds, dim=1)


def _tv_adaptor(w, p, x, x_mask, sty, sty_mask, time_tok):
    """TVAdaptor.forward, DEX-TTS/model/ref_encoder.py:154-179.
    x (B,C,H,W); sty (B,C,Ts); sty_mask (B,Ts) in {0,1}; time_tok (B,C)."""
    b, c, hh, ww = x.shape
    toks = torch.cat([time_tok.unsqueeze(-1), sty], dim=-1).transpose(1, 2)           # (B, Ts+1, C)
    kmask = torch.cat([torch.ones(b, 1, dtype=sty_mask.dtype), sty_mask], dim=-1)      # time token always visible
    q = F.linear(_instance_norm_2d(x).permute(0, 2, 3, 1), w[p + ".w_q.weight"])     # (B,H,W,C)
    k = F.linear(toks, w[p + ".w_k.weight"]).unsqueeze(1)                              # (B,1,Ts+1,C)
    v = F.linear(toks, w[p + ".w_v.weight"]).unsqueeze(1)
    att = torch.matmul(q / (c ** 0.5), k.transpose(-1, -2))                            # (B,H,W,Ts+1)
    att = att.masked_fill(kmask[:, None, None, :] == 0, -1e4)
    att = att.softmax(dim=-1)
    out = F.linear(torch.matmul(att, v), w[p + ".linear.weight"]).permute(0, 3, 1, 2)
    return (x + out) * x_mask


def _sap(w, p, rows, time_tok):
    """SelfAttentionPooling, DEX-TTS/model/ref_encoder.py:246-253.  rows (B,L,C), time_tok (B,1,C) -> (B,C)."""
    z = torch.cat([time_tok, rows], dim=1)
    a = F.linear(z, w[p + ".W.weight"], w[p + ".W.bias"]).squeeze(-1)
    a = a.softmax(dim=-1).unsqueeze(-1)
    return (z * a).sum(dim=1)


def _tiv_adaptor(w, p, x, ref_mean, ref_std, time_tok):
    """TIVAdaptor.forward, DEX-TTS/model/ref_encoder.py:264-273 (AdaIN; output is NOT re-masked)."""
    m = _sap(w, p + ".mean_sap", ref_mean, time_tok)[:, :, None, None]
    s = _sap(w, p + ".std_sap", ref_std, time_tok)[:, :, None, None]
    return _instance_norm_2d(x) * s + m


def _modulate(x, shift, scale):
    """DEX-TTS/model/dit.py:72-73."""
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def _dit_block(w, p, x, c, heads):
    """DiTBlock.forward, DEX-TTS/model/dit.py:280-284, with timm Attention (qkv bias, no mask) and Mlp (exact GELU)."""
    bsz, n, d = x.shape
    hd = d // heads
    mod = F.linear(F.silu(c), w[p + ".adaLN_modulation.1.weight"], w[p + ".adaLN_modulation.1.bias"])
    sh_a, sc_a, g_a, sh_m, sc_m, g_m = mod.chunk(6, dim=1)
    h = _modulate(F.layer_norm(x, (d,), eps=1e-6), sh_a, sc_a)
    qkv = _lin(w, p + ".attn.qkv", h).reshape(bsz, n, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    att = ((q * hd ** -0.5) @ k.transpose(-2, -1)).softmax(dim=-1)
    a = (att @ v).transpose(1, 2).reshape(bsz, n, d)
    x = x + g_a.unsqueeze(1) * _lin(w, p + ".attn.proj", a)
    h = _modulate(F.layer_norm(x, (d,), eps=1e-6), sh_m, sc_m)
    h = _lin(w, p + ".mlp.fc2", F.gelu(_lin(w, p + ".mlp.fc1", h)))
    return x + g_m.unsqueeze(1) * h


def dit_grid(cfg, w_in, h_in=None):
    """Token grid (Fq, Wq, Wpad) for a bottleneck image of width w_in.  DEX-TTS/model/dit.py:434-450,31-54."""
    p, s = cfg["patch"], cfg["stride"]
    if h_in is None:
        h_in = cfg["n_feats"] // 2
    wp = w_in if w_in % p == 0 else w_in + (p - w_in % p)
    fq = (h_in + 2 * (p // 2) - p) // s + 1
    wq = (wp + 2 * (p // 2) - p) // s + 1
    return fq, wq, wp


def _dit(w, p, cfg, x, mask, t, taps=None):
    """DiTMask.forward (inference branch), DEX-TTS/model/dit.py:479-519."""
    ps, st, hid = cfg["patch"], cfg["stride"], cfg["hidden"]
    x_len = x.shape[-1]
    if x_len % ps != 0:                                                   # pad to a multiple of PATCH size (:436-439)
        x = F.pad(x, (0, ps - x_len % ps))
    c_in = x.shape[1]
    e = F.conv2d(x, w[p + ".x_embedder.proj.0.weight"], w[p + ".x_embedder.proj.0.bias"],
                 stride=st, padding=ps // 2, groups=c_in)
    e = F.conv2d(F.silu(e), w[p + ".x_embedder.proj.2.weight"], w[p + ".x_embedder.proj.2.bias"])
    kp = cfg["conv_pos"]
    pe = F.conv2d(e, w[p + ".pos_conv.0.weight"], w[p + ".pos_conv.0.bias"], padding=kp // 2,
                  groups=cfg["conv_pos_groups"])
    if kp % 2 == 0:                                                       # SamePad, dit.py:122-133
        pe = pe[:, :, :-1, :-1]
    pe = F.gelu(pe).mean(dim=2, keepdim=True)
    e = e + pe[:, :, :, :e.shape[-1]]
    e = e + w[p + ".freq_new_pos_embed"]
    if taps is not None:
        taps["dit_tokens"] = e
    tok = e.flatten(2).transpose(1, 2)
    temb = timestep_embedding(t, 256)                                     # frequency_embedding_size=256 (:223)
    c = _lin(w, p + ".t_embedder.mlp.2", F.silu(_lin(w, p + ".t_embedder.mlp.0", temb)))
    for i in range(cfg["depth"]):
        tok = _dit_block(w, f"{p}.blocks.{i}", tok, c, cfg["heads"])
    mod = F.linear(F.silu(c), w[p + ".final_layer.adaLN_modulation.1.weight"],
                   w[p + ".final_layer.adaLN_modulation.1.bias"])
    sh, sc = mod.chunk(2, dim=1)
    tok = _lin(w, p + ".final_layer.linear", _modulate(F.layer_norm(tok, (hid,), eps=1e-6), sh, sc))
    # unpatchify 'B (h w) (p1 p2 C) -> B C (h p1) (w p2)', p = sqrt(out_dim / C), h = input_size[0] // p (:452-457)
    pp = int((tok.shape[2] // c_in) ** 0.5)
    hq = int((cfg["n_feats"] // 2) // pp)
    bsz, n, _ = tok.shape
    wq = n // hq
    img = tok.reshape(bsz, hq, wq, pp, pp, c_in).permute(0, 5, 1, 3, 2, 4).reshape(bsz, c_in, hq * pp, wq * pp)
    return img[..., :x_len] * mask


# ----------------------------------------------------------------------------------------------
# denoiser, preconditioning, sampler
# ----------------------------------------------------------------------------------------------

def denoiser(w, cfg, x, mask, mu, t, cond=None, prefix="denoise_fn", taps=None):
    """DiffusionDenoiser.forward.  x, mu (B,80,T); mask (B,1,T); t (B,) = c_noise.
    cond: DEX = dict(sty (B,128,Ts), sty_lengths (B,), ref_skips [6 x (B,128,Tr)]); multi-speaker GeDEX = dict(spk (B, spk_emb_dim)).
    DEX-TTS/model/diffusion.py:190-236, GeDEX-TTS/model/diffusion.py:168-207."""
    p = prefix
    dex = cfg["variant"] == "dex"
    dim = cfg["dim"]
    t_init = sinusoidal_emb(t, dim, cfg["pe_scale"])
    t_unet = _lin(w, p + ".mlp.2", mish(_lin(w, p + ".mlp.0", t_init)))
    if dex:
        t_adap = _lin(w, p + ".mlp_adap.2", mish(_lin(w, p + ".mlp_adap.0", t_init)))          # (B, 2*dim)
        t_sty = _lin(w, p + ".mlp_adap_sty.2", mish(_lin(w, p + ".mlp_adap_sty.0", t_init)))
        sty_mask = sequence_mask(cond["sty_lengths"], cond["sty"].shape[2]).to(x.dtype)
        ref_mean, ref_std = ref_stats(cond["ref_skips"])
    if not dex and cfg.get("n_spks", 1) > 1:
        # GeDEX-TTS/model/diffusion.py:170-175: s = spk_mlp(spk) broadcast over time is the third input channel
        s_ = _lin(w, p + ".spk_mlp.2", mish(_lin(w, p + ".spk_mlp.0", cond["spk"])))
        h = torch.stack([mu, x, s_.unsqueeze(-1).repeat(1, 1, x.shape[-1])], 1)
    else:
        h = torch.stack([mu, x], 1)
    m0 = mask.unsqueeze(1)                                                                     # (B,1,1,T)
    # level 0
    h = _resnet(w, p + ".downs.0.0", h, m0, t_unet)
    if taps is not None:
        taps["d00"] = h
    h = _resnet(w, p + ".downs.0.1", h, m0, t_unet)
    if taps is not None:
        taps["d01"] = h
    h = _linear_attention(w, p + ".downs.0.2", h)
    h = F.conv2d(h * m0, w[p + ".downs.0.3.conv.weight"], w[p + ".downs.0.3.conv.bias"], stride=2, padding=1)
    m1 = m0[:, :, :, ::2]
    # level 1
    h = _resnet(w, p + ".downs.1.0", h, m1, t_unet)
    h = _resnet(w, p + ".downs.1.1", h, m1, t_unet)
    h = _linear_attention(w, p + ".downs.1.2", h)
    skip = h
    h = h * m1
    if taps is not None:
        taps["down_out"] = h
        taps["skip"] = h
    if dex:
        h = _tv_adaptor(w, p + ".tv_adaptor", h, m1, cond["sty"], sty_mask, t_sty)
        if taps is not None:
            taps["tv_out"] = h
        h = _tiv_adaptor(w, p + ".tiv_adaptor", h, ref_mean, ref_std, t_adap.unsqueeze(1))
        if taps is not None:
            taps["tiv_out"] = h
    h = _dit(w, p + ".vit", cfg, h, m1, t, taps=taps)
    if taps is not None:
        taps["dit_out"] = h
    # up
    h = torch.cat((h, skip), dim=1)
    h = _resnet(w, p + ".ups.0.0", h, m1, t_unet)
    if taps is not None:
        taps["u00"] = h
    h = _resnet(w, p + ".ups.0.1", h, m1, t_unet)
    if taps is not None:
        taps["u01"] = h
    h = _linear_attention(w, p + ".ups.0.2", h)
    h = F.conv_transpose2d(h * m1, w[p + ".ups.0.3.conv.weight"], w[p + ".ups.0.3.conv.bias"], stride=2, padding=1)
    if taps is not None:
        taps["up_out"] = h
    h = _block(w, p + ".final_block", h, m0)
    out = F.conv2d(h * m0, w[p + ".final_conv.weight"], w[p + ".final_conv.bias"])
    return (out * m0).squeeze(1)


def edm_precond(w, cfg, x, sigma, mask, mu, cond=None, sigma_data=0.5, taps=None):
    """EDMPrecond.forward, DEX-TTS/model/edm.py:88-98.  sigma: 0-dim or (B,) tensor."""
    sigma = sigma.reshape(-1).expand(x.shape[0]).reshape(-1, 1, 1)
    c_skip = sigma_data ** 2 / (sigma ** 2 + sigma_data ** 2)
    c_out = sigma * sigma_data / (sigma ** 2 + sigma_data ** 2).sqrt()
    c_in = 1 / (sigma_data ** 2 + sigma ** 2).sqrt()
    c_noise = sigma.log() / 4
    f_x = denoiser(w, cfg, c_in * x, mask, mu, c_noise.flatten(), cond=cond, taps=taps)
    return c_skip * x + c_out * f_x


def edm_loss(w, cfg, x0, mask, mu, rnd_normal, noise, cond=None, P_mean=-1.2, P_std=1.2, sigma_data=0.5, n_feats=80):
    """EDMLoss.forward with loss_type 'base', DEX-TTS/model/edm.py:31-38,64-68 (the training branch of Diffusion.forward,
    diffusion.py:252-254), with its two Gaussian draws injected: rnd_normal (B,1,1) -> the per-utterance noise
    level, noise (B,80,T) -> the prior-shifted perturbation (noise + mu) * sigma.  Forward value only; no CUDA side."""
    sigma = (rnd_normal * P_std + P_mean).exp()                                                       # :34
    weight = (sigma ** 2 + sigma_data ** 2) / (sigma * sigma_data) ** 2                               # :38
    n = (noise + mu) * sigma                                                                          # :64
    d_yn = edm_precond(w, cfg, x0 + n, sigma.reshape(-1), mask, mu, cond=cond, sigma_data=sigma_data)  # :65
    return torch.sum(weight * ((d_yn - x0) ** 2)) / torch.sum(mask * n_feats)                         # :66


def sigma_schedule(num_steps, sigma_min=0.002, sigma_max=80.0, rho=7, dtype=torch.float32):
    """EDM discretisation in fp32 as the reference computes it, plus the trailing 0.  DEX-TTS/model/edm.py:135-152,179-180."""
    idx = torch.arange(num_steps)
    s = (sigma_max ** (1 / rho) + idx / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    s = s.to(dtype)
    return torch.cat([s, torch.zeros_like(s[:1])])


def sampler(w, cfg, latents, mask, mu, num_steps, cond=None, trace=None):
    """ablation_sampler(solver='euler', discretization='edm', schedule='linear', scaling='none'), S_churn=0.
    DEX-TTS/model/edm.py:183-209: x <- latents*sigma_0; N x { D = net(x, sigma_i); x <- x + (sigma_{i+1}-sigma_i)*(x-D)/sigma_i }."""
    ts = sigma_schedule(num_steps, dtype=latents.dtype)
    x = latents * ts[0]
    for i in range(num_steps):
        t_cur, t_next = ts[i], ts[i + 1]
        den = edm_precond(w, cfg, x, t_cur, mask, mu, cond=cond)
        d_cur = (1 / t_cur) * x - 1 / t_cur * den          # same operation order as edm.py:197
        x = x + (t_next - t_cur) * d_cur
        if trace is not None:
            trace.append(x.clone())
    return x


def reverse_diffusion(w, cfg, z, mask, mu, num_steps, temperature=1.0, cond=None, trace=None):
    """Diffusion.forward(infer=True) with the Gaussian draw ``z`` injected (the reference draws it on-device).
    DEX-TTS/model/diffusion.py:255-259:  x = z / temperature + mu ; x = sampler(x, ...)."""
    x = z / temperature + mu
    return sampler(w, cfg, x, mask, mu, num_steps, cond=cond, trace=trace)


# ----------------------------------------------------------------------------------------------
# pre-loop stage: TIV encoder (produces the six `ref` skip tensors of the loop's TIVAdaptor)
# ----------------------------------------------------------------------------------------------

def _basic_conv_bn(w, p, x, relu=True, norm=True):
    """BasicConv(norm_type='bn') in eval mode, DEX-TTS/model/base.py:33-63: conv1d(k=3, pad 1, no bias) -> BatchNorm1d on
    running statistics (eps 1e-5) -> ReLU."""
    x = F.conv1d(x, w[p + ".conv.weight"], None, padding=1)
    if norm:
        x = F.batch_norm(x, w[p + ".bn.running_mean"], w[p + ".bn.running_var"], w[p + ".bn.weight"], w[p + ".bn.bias"],
                         training=False, momentum=0.01, eps=1e-5)
    return torch.relu(x) if relu else x


def _instance_norm_1d(x, eps=1e-5):
    """InstanceNorm1D.forward, DEX-TTS/model/base.py:66-93: statistics over ALL frames (x_lengths ignored), unbiased variance."""
    mean = x.mean(-1, keepdim=True)
    std = (x.var(-1, keepdim=True) + eps).sqrt()
    return (x - mean) / std


def tiv_encoder(w, ref, mask, num_layer=6, prefix="tiv_encoder"):
    """TIVEncoder.forward(x, mask), DEX-TTS/model/ref_encoder.py:95-107 (eval).  ref (B,80,T) or (B,1,80,T), mask (B,1,T)
    -> (out (B,c_out,T), [num_layer x (B,c_h,T)])."""
    if ref.dim() == 4:
        ref = ref.squeeze(1)
    x = _basic_conv_bn(w, prefix + ".in_conv", ref * mask) * mask                                   # :97
    skips = []
    for i in range(num_layer):
        p = f"{prefix}.conv_blocks.{i}.conv_block"
        xin = x * mask
        h = _basic_conv_bn(w, p + ".0", xin)                                                         # :61
        x = (xin + _basic_conv_bn(w, p + ".1", h, relu=False, norm=False)) * mask                    # :62,66,101
        skips.append(x)
        x = _instance_norm_1d(x)                                                                     # :103
    out = _basic_conv_bn(w, prefix + ".out_conv", x * mask) * mask                                   # :104
    return out, skips


# ----------------------------------------------------------------------------------------------
# pre-loop stage: TV encoder (produces z_dec, from which DeXTTS.forward builds the loop's `sty`)
# ----------------------------------------------------------------------------------------------

def _basic_conv_ln(w, p, x):
    """BasicConv(norm_type='ln'), DEX-TTS/model/base.py:33-63: conv1d(k=3, pad 1, no bias) -> ReLU -> nn.LayerNorm over channels
    (eps 1e-5, affine)."""
    x = torch.relu(F.conv1d(x, w[p + ".conv.weight"], None, padding=1))
    c = x.shape[1]
    return F.layer_norm(x.transpose(1, 2), (c,), w[p + ".ln.weight"], w[p + ".ln.bias"], 1e-5).transpose(1, 2)


def _channel_layer_norm(w, p, x, eps=1e-4):
    """model.base.LayerNorm, DEX-TTS/model/base.py:139-159 (channel axis of (B,C,T), biased variance, rsqrt, eps 1e-4)."""
    mean = torch.mean(x, 1, keepdim=True)
    variance = torch.mean((x - mean) ** 2, 1, keepdim=True)
    x = (x - mean) * torch.rsqrt(variance + eps)
    return x * w[p + ".gamma"].view(1, -1, 1) + w[p + ".beta"].view(1, -1, 1)


def _projection(w, p, x, mask):
    """Projection.forward (eval: dropout = identity), DEX-TTS/model/ref_encoder.py:24-34."""
    x = F.conv1d(x * mask, w[p + ".conv_1.weight"], w[p + ".conv_1.bias"], padding=1)
    x = _channel_layer_norm(w, p + ".norm_1", torch.relu(x))
    x = F.conv1d(x * mask, w[p + ".conv_2.weight"], w[p + ".conv_2.bias"], padding=1)
    x = _channel_layer_norm(w, p + ".norm_2", torch.relu(x))
    x = F.conv1d(x * mask, w[p + ".proj.weight"], w[p + ".proj.bias"])
    return x * mask


def vq_indices(w, p, x_flat):
    """Nearest-code search of VQEmbeddingEMA.forward, DEX-TTS/model/ref_encoder.py:206-211."""
    emb = w[p + ".embedding"]
    distances = torch.addmm(torch.sum(emb ** 2, dim=1) + torch.sum(x_flat ** 2, dim=1, keepdim=True), x_flat, emb.t(),
                            alpha=-2.0, beta=1.0)
    return torch.argmin(distances.float(), dim=-1), distances


def _vq(w, p, x, mask, commit_w=0.25):
    """VQEmbeddingEMA.forward in eval mode, DEX-TTS/model/ref_encoder.py:199-235.  x (B,T,D), mask (B,1,T)."""
    m = mask.transpose(1, 2)
    x = x * m
    D = x.shape[-1]
    idx, _ = vq_indices(w, p, x.reshape(-1, D))
    quantized = F.embedding(idx, w[p + ".embedding"]).view_as(x)
    loss = commit_w * torch.sum(((x * m) - (quantized * m)) ** 2) / (torch.sum(m) * D)
    quantized = x + (quantized - x)
    return quantized * m, loss, idx.view(x.shape[:2])


def tv_encoder(w, sty, mask, num_layer=6, commit_w=0.25, prefix="tv_encoder", return_indices=False):
    """TVEncoder.forward(x, mask), DEX-TTS/model/ref_encoder.py:124-140 (eval).  sty (B,80,T) or (B,1,80,T), mask (B,1,T)
    -> (z_beforeVQ (B,c_out,T), z_dec (B,c_out_g,T), vq_loss)."""
    if sty.dim() == 4:
        sty = sty.squeeze(1)
    x = _basic_conv_ln(w, prefix + ".in_conv", sty * mask) * mask                                     # :128
    for i in range(num_layer):
        p = f"{prefix}.conv_blocks.{i}.conv_block"
        xin = x * mask
        h = _basic_conv_ln(w, p + ".0", xin)
        x = (xin + F.conv1d(h, w[p + ".1.conv.weight"], None, padding=1)) * mask                     # :73-79,132
    z_before = F.conv1d(x * mask, w[prefix + ".out_conv.conv.weight"], None, padding=1) * mask        # :134
    z, loss, idx = _vq(w, prefix + ".vq", z_before.transpose(1, 2), mask, commit_w)                   # :135
    z_dec = _projection(w, prefix + ".proj_0", z.transpose(1, 2), mask)                               # :137-138
    z_dec = _basic_conv_bn(w, prefix + ".proj_1", z_dec * mask) * mask                                # :139
    if return_indices:
        return z_before, z_dec, loss, idx
    return z_before, z_dec, loss


# ----------------------------------------------------------------------------------------------
# pre-loop stage: LF0 encoder and the style fusion of DeXTTS.forward
# ----------------------------------------------------------------------------------------------

def _gru_direction(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of one nn.GRU layer (PyTorch gate order r, z, n; h0 = 0).  x (B,T,I) -> (B,T,H)."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    gi = x @ w_ih.t() + b_ih                                  # (B,T,3H)
    h = x.new_zeros(B, H)
    out = x.new_zeros(B, T, H)
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        gh = h @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[:, t, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, t, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, t, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * h
        out[:, t] = h
    return out


def _bigru(w, p, x, num_layer):
    """nn.GRU(c_h, c_h // 2, num_layer, batch_first=True, bidirectional=True) in eval mode, DEX-TTS/model/ref_encoder.py:41,50:
    no packing -- both directions run over all T frames, padding included."""
    for l in range(num_layer):
        outs = []
        for sfx, rev in (("", False), ("_reverse", True)):
            outs.append(_gru_direction(x, w[f"{p}.weight_ih_l{l}{sfx}"], w[f"{p}.weight_hh_l{l}{sfx}"],
                                       w[f"{p}.bias_ih_l{l}{sfx}"], w[f"{p}.bias_hh_l{l}{sfx}"], rev))
        x = torch.cat(outs, dim=-1)
    return x


def lf0_encoder(w, lf0, mask, num_layer=2, prefix="lf0_encoder"):
    """LF0Encoder.forward(lf0, mask), DEX-TTS/model/ref_encoder.py:46-56 (eval).  lf0 (B,T), mask (B,1,T)
    -> (lf0_enc (B,c_out,T), lf0_dec (B,c_out_g,T))."""
    x = lf0.unsqueeze(1)
    x = _basic_conv_ln(w, prefix + ".in_conv", x * mask) * mask                                       # :49
    x = _bigru(w, prefix + ".rnn_layer", x.transpose(1, 2), num_layer)                               # :50
    x = _basic_conv_ln(w, prefix + ".out_conv", x.transpose(1, 2) * mask) * mask                     # :51
    return x, _projection(w, prefix + ".proj", x, mask)                                              # :53-54


def style_fusion(w, sty_enc_tv, sty_dec_tv, lf0_enc, lf0_dec, sty_mask, lf0_mask):
    """DeXTTS.forward, DEX-TTS/model/tts.py:45-49: masked time means of the LF0 branch added to the TV branch, then conv_sty.
    -> (sty_enc (B,c_out) for the text encoder, sty_dec (B,2*dim,Ts) = the loop's `sty`)."""
    sty_enc = (sty_enc_tv.sum(dim=-1) / sty_mask.sum(dim=-1)) + (lf0_enc.sum(dim=-1) / lf0_mask.sum(dim=-1))
    sty_dec = sty_dec_tv + (lf0_dec.sum(dim=-1) / lf0_mask.sum(dim=-1)).unsqueeze(-1)
    sty_dec = F.conv1d(sty_dec, w["conv_sty.weight"], w["conv_sty.bias"])
    return sty_enc, sty_dec


def align_durations(logw, x_mask, mu_x, length_scale=1.0):
    """Duration / alignment glue of DeXTTS.forward, DEX-TTS/model/tts.py:55-68 (GeDEX-TTS/model/tts.py:37-50), with
    sequence_mask / fix_len_compatibility / generate_path of DEX-TTS/model/utils.py:6-39 written out as index arithmetic
    (pinned against those functions themselves by tests/golden/align_*.npz, oracle/make_golden_align.py).
    logw, x_mask (B,1,Tx), mu_x (B,F,Tx) -> (mu_y (B,F,Ty_), y_mask (B,1,Ty_), attn (B,1,Tx,Ty_), y_lengths (B,), y_max_length)."""
    w_ceil = torch.ceil(torch.exp(logw) * x_mask) * length_scale                                      # tts.py:55-56
    y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()                                  # :57
    y_max = int(y_lengths.max())                                                                      # :58
    Ty = y_max
    while Ty % 4:                                                                                     # utils.py:13-17
        Ty += 1
    t = torch.arange(Ty, dtype=torch.float32)
    y_mask = (t[None, :] < y_lengths[:, None].float()).float().unsqueeze(1)                           # :62 (lengths < 2^24: exact)
    cum = torch.cumsum(w_ceil.squeeze(1), 1)                                                          # utils.py:30
    start = torch.cat([torch.zeros_like(cum[:, :1]), cum[:, :-1]], 1)
    # generate_path: row i of `t < cum[i]` minus row i-1 (utils.py:34-37)  ==  cum[i-1] <= t < cum[i] for a non-decreasing cum
    path = ((t[None, None, :] < cum[:, :, None]) & (t[None, None, :] >= start[:, :, None])).float()
    attn = path * (x_mask.squeeze(1)[:, :, None] * y_mask)                                            # utils.py:38, tts.py:63
    mu_y = torch.matmul(attn.transpose(1, 2), mu_x.transpose(1, 2)).transpose(1, 2)                   # :67-68
    return mu_y, y_mask, attn.unsqueeze(1), y_lengths, y_max


def decoder_weights(state_dict, dtype=torch.float32):
    """Keep the ``denoise_fn.*`` view of a Diffusion state_dict (the ``precond_model.model.*`` keys alias the same
    tensors, DEX-TTS/model/diffusion.py:242-243)."""
    return {k: v.detach().to(dtype) for k, v in state_dict.items() if k.startswith("denoise_fn.")}


# ----------------------------------------------------------------------------------------------
# algorithmic work per unit -- used by bench.py for the roofline line
# ----------------------------------------------------------------------------------------------

def flops_per_sample_step(cfg, T, Ts=259):
    """Algorithmic FLOPs (2*MAC) of one denoiser call for one utterance; closed form of SURVEY.md §8(d)."""
    p, s = cfg["patch"], cfg["stride"]
    W = T // 2
    fq, wq, _ = dit_grid(cfg, W)
    n = fq * wq
    f_unet = 70.994e6 * T
    f_dit = 2 * (n * (p * p * 128 + 128 * 256) + (fq + 1) * (wq + 1) * 2097152 + 4 * n * 524288
                 + 4 * 2 * n * n * 256 + n * 256 * s * s * 128)
    f_tv = 2 * (20 * T * (32768 + 256 * (Ts + 1)) + 32768 * (Ts + 1)) if cfg["variant"] == "dex" else 0
    return f_unet + f_dit + f_tv
