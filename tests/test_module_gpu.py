"""The drop-in modules (dexb200.model.Diffusion / GeDiffusion) called the way DeXTTS.forward / GeDEXTTS.forward call the
reference decoder (DEX-TTS/model/tts.py:71, GeDEX-TTS/model/tts.py:53), against the CPU oracle fed with the same noise."""
import pytest
import torch

import dex_oracle as O
from dexb200.manifest import DecoderCfg
from dexb200.synth import synth_decoder_weights, synth_inputs
from parity import REL_TOL, per_bin_violation

pytestmark = pytest.mark.gpu


def make_module(variant):
    from dexb200.model import Diffusion, GeDiffusion
    dit = dict(patch_size=3 if variant == "dex" else 7, stride_size=2 if variant == "dex" else 4, hidden_size=256, depth=4,
               num_heads=2, mlp_ratio=2, conv_pos=16, conv_pos_groups=8, in_channels=3, out_channels=1)
    cls = Diffusion if variant == "dex" else GeDiffusion
    m = cls(n_feats=80, dim=64, dit_cfg=dit, dim_mults=[1, 2], model_type="dit", n_spks=0 if variant == "dex" else 1, pe_scale=1000)
    cfg = DecoderCfg.make(variant)
    w = synth_decoder_weights(cfg, seed=100, live=True)
    sd = dict(w)
    sd.update({k.replace("denoise_fn.", "precond_model.model."): v for k, v in w.items()})
    m.load_state_dict(sd, strict=True)                       # reference checkpoint layout: both prefixes
    return m.cuda().eval(), cfg, w


@pytest.mark.parametrize("variant", ["dex", "gedex"])
def test_module_forward_matches_oracle(variant):
    m, cfg, w = make_module(variant)
    B, T, Ts, Tr, steps, temperature = 2, 64, 37, 29, 6, 1.5
    inp = synth_inputs(cfg, B, T, Ts=Ts, seed=5, ragged=True)
    mu, mask = inp["mu"].cuda(), inp["mask"].cuda()
    if variant == "dex":
        g = torch.Generator().manual_seed(9)
        refs = [torch.randn(B, 128, Tr, generator=g) for _ in range(6)]          # Tr != Ts on purpose
        ref_lengths = torch.full((B,), Tr, dtype=torch.long)
        args = (mu, mask, mu, [r.cuda() for r in refs], ref_lengths.cuda(), inp["sty"].cuda(), inp["sty_lengths"].cuda())
        cond = dict(sty=inp["sty"], sty_lengths=inp["sty_lengths"], ref_skips=refs)
    else:
        args, cond = (mu, mask, mu), None
    torch.manual_seed(1234)
    y = m(*args, n_timesteps=steps, temperature=temperature, spk=None, infer=True).cpu()
    torch.manual_seed(1234)
    z = torch.randn((B, 80, T), device="cuda").cpu()          # the draw Diffusion.forward makes (diffusion.py:256-257)
    with torch.no_grad():
        y_ref = O.reverse_diffusion(w, O.make_cfg(variant), z, inp["mask"], inp["mu"], steps, temperature=temperature, cond=cond)
    assert y.shape == (B, 80, T)
    assert per_bin_violation(y, y_ref) < REL_TOL
    # parameters changed in place are picked up (weights are re-packed): zero the Rezero gates -> different output
    with torch.no_grad():
        for n_, p_ in m.denoise_fn.named_parameters():
            if n_.endswith("fn.g"):
                p_.zero_()
    torch.manual_seed(1234)
    y2 = m(*args, n_timesteps=steps, temperature=temperature, spk=None, infer=True).cpu()
    assert not torch.equal(y2, y)
    w2 = {k: (torch.zeros_like(v) if k.endswith("fn.g") else v) for k, v in w.items()}
    with torch.no_grad():
        y2_ref = O.reverse_diffusion(w2, O.make_cfg(variant), z, inp["mask"], inp["mu"], steps, temperature=temperature, cond=cond)
    assert per_bin_violation(y2, y2_ref) < REL_TOL
