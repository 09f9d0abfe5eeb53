"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only:   python oracle/make_golden.py
The GPU box has no /root/reference; it replays the committed fixtures.

For every case the script
  1. builds the reference ``Diffusion`` (DEX-TTS/model/diffusion.py:238 / GeDEX-TTS/model/diffusion.py:209),
  2. loads the deterministic synthetic weights of ``dexb200.synth`` with ``load_state_dict(strict=True)``
     (this also pins the manifest's key/shape compatibility with upstream checkpoints),
  3. runs ``decoder(..., infer=True)`` with ``torch.randn`` patched to return the case's CPU noise,
  4. stores inputs' seeds, the reference output and a few intermediate activations of the first net call.
Batched DEX (B>1) needs sigma broadcast to (B,) (reference bug): done by wrapping
``EDMPrecond.forward`` -- per-sample arithmetic is unchanged (case ``dex_b1`` runs the un-wrapped reference).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "dex-tts_b200"))

import ref_loader                                    # noqa: E402
from dexb200.manifest import DecoderCfg              # noqa: E402
from dexb200.synth import synth_decoder_weights, synth_inputs   # noqa: E402

CASES = [
    # name,        variant, B, T,  Ts, steps, ragged, live, seed
    ("gedex_b1",   "gedex", 1, 48, 0,  4, False, True, 11),
    ("gedex_b2r",  "gedex", 2, 52, 0,  3, True,  True, 12),
    ("gedex_zero", "gedex", 1, 40, 0,  3, False, False, 13),   # reference default-style init (dead DiT branch)
    ("dex_b1",     "dex",   1, 44, 23, 4, False, True, 21),
    ("dex_b2r",    "dex",   2, 48, 19, 3, True,  True, 22),
    ("gedex_spk_b2r", "gedex", 2, 44, 0, 3, True, True, 14, 4),    # multi-speaker GeDEX-TTS: n_spks = 4 (third input channel)
    # DEX-TTS/config/LibriTTS/base.yaml:65,77: decoder dim 128, DiT hidden 384 (head dim 192).  Oracle-only fixture: the CUDA engine
    # instantiates dim 64 / hidden 256 so far, so the name keeps it out of the GPU tests' dex_* / gedex_* globs.
    ("libri_dex_b1", "dex", 1, 44, 23, 3, False, True, 31, None, dict(dim=128, hidden=384)),
    # the benchmarked mel length (BASELINE.json configs 2-4: 80x512, style length 259): different code runs there -- four halo tiles
    # per row, the attention tail split, split-KV linear-attention contexts.  Taps are additionally strided along time.
    ("dex_t512_b1", "dex",  1, 512, 259, 3, False, True, 41),
    ("gedex_t512_b1", "gedex", 1, 512, 0, 3, False, True, 42),
]
TEMPERATURE = 1.5
TAP_STRIDE = 16
TAP_WSTRIDE = 8          # additional stride along time for the T >= 256 cases


def ref_cfgs(cfg):
    dec = dict(dim=cfg.dim, pe_scale=cfg.pe_scale, dim_mults=[1, 2], model_type="dit", precond="edm", loss_type="base")
    dit = dict(in_channels=3, patch_size=cfg.patch, stride_size=cfg.stride, overlap=True, hidden_size=cfg.hidden,
               depth=cfg.depth, num_heads=cfg.heads, mlp_ratio=cfg.mlp_ratio, out_channels=1, conv_pos=cfg.conv_pos,
               conv_pos_groups=cfg.conv_pos_groups, use_decoder=False, mask_type="time_random")
    return dec, dit


def run_case(name, variant, B, T, Ts, steps, ragged, live, seed, n_spks=None, dims=None):
    cfg = DecoderCfg.make(variant, n_spks=n_spks, **(dims or {}))
    dec_cfg, dit_cfg = ref_cfgs(cfg)
    dec, mod = ref_loader.build_reference_decoder(variant, dec_cfg, dit_cfg, n_spks=n_spks)
    w = synth_decoder_weights(cfg, seed=100, live=live)
    sd = dict(w)
    sd.update({k.replace("denoise_fn.", "precond_model.model."): v for k, v in w.items()})
    dec.load_state_dict(sd, strict=True)
    inp = synth_inputs(cfg, B, T, Ts=max(Ts, 1), seed=seed, ragged=ragged)

    if B > 1:
        orig = mod.EDMPrecond.forward

        def fwd(self, x, sigma, *a, **k):
            return orig(self, x, sigma.reshape(-1).expand(x.shape[0]), *a, **k)
        mod.EDMPrecond.forward = fwd

    taps = {}
    dn = dec.denoise_fn
    hooks = []
    first = {"done": False}

    def tap(key):
        def h(m, i, o):
            if key not in taps:
                taps[key] = o.detach().clone()
        return h
    hooks.append(dn.downs[1][2].register_forward_hook(tap("skip")))
    if variant == "dex":
        hooks.append(dn.tv_adaptor.register_forward_hook(tap("tv_out")))
        hooks.append(dn.tiv_adaptor.register_forward_hook(tap("tiv_out")))
    hooks.append(dn.vit.register_forward_hook(tap("dit_out")))
    hooks.append(dn.ups[0][3].register_forward_hook(tap("up_out")))
    hooks.append(dn.register_forward_hook(tap("f_x0")))

    real_randn = torch.randn
    torch.randn = lambda *a, **k: inp["z"].clone()
    try:
        with torch.no_grad():
            if variant == "dex":
                y = dec(inp["mu"], inp["mask"], inp["mu"], inp["ref_skips"], inp["ref_lengths"], inp["sty"],
                        inp["sty_lengths"], n_timesteps=steps, infer=True, temperature=TEMPERATURE)
            else:
                y = dec(inp["mu"], inp["mask"], inp["mu"], n_timesteps=steps, spk=inp.get("spk"), infer=True,
                        temperature=TEMPERATURE)
    finally:
        torch.randn = real_randn
        for h in hooks:
            h.remove()
    out = dict(y=y.numpy(), meta=np.array([B, T, Ts, steps, int(ragged), int(live), seed] + ([int(n_spks)] if n_spks else []), dtype=np.int64),
               variant=np.array(variant), temperature=np.array(TEMPERATURE, dtype=np.float32),
               dims=np.array([cfg.dim, cfg.hidden], dtype=np.int64))
    for k, v in taps.items():               # intermediates: channel-strided subsample keeps the fixtures small
        a = v.numpy().astype(np.float32)
        wst = TAP_WSTRIDE if T >= 256 else 1
        out["tap_" + k] = a if a.ndim == 3 else a[:, ::TAP_STRIDE, :, ::wst]
    out["tap_wstride"] = np.array(TAP_WSTRIDE if T >= 256 else 1, dtype=np.int64)
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: y {tuple(y.shape)} |y|max {float(y.abs().max()):.4f} -> {os.path.relpath(path, ROOT)} "
          f"({os.path.getsize(path)/1024:.0f} KiB)")


if __name__ == "__main__":
    torch.set_num_threads(8)
    only = set(sys.argv[1:])                         # optional: names of the cases to (re)generate
    for c in CASES:
        if not only or c[0] in only:
            run_case(*c)
