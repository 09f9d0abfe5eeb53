"""GPU parity of the LF0 encoder and the style fusion (dexb_lf0_*, dexb_style_fuse through the C ABI) against the reference
fixtures and the CPU oracle, and the whole pre-loop stage chained into the reverse diffusion the way DeXTTS.forward orders it."""
import glob
import os

import numpy as np
import pytest
import torch

import dex_oracle as O
from dexb200.manifest import DecoderCfg
from dexb200.synth import (synth_conv_sty_weights, synth_decoder_weights, synth_inputs, synth_lf0, synth_lf0_weights, synth_ref_mel,
                           synth_tiv_weights, synth_tv_weights)
from parity import REL_TOL, per_bin_violation, tensor_rel_err

pytestmark = pytest.mark.gpu

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "lf0_*.npz")))
TOL = 1e-3             # the path tolerance (north_star), as max |a - b| / RMS(reference tensor)


def make_modules():
    from dexb200.model import LF0Encoder, TIVEncoder, TVEncoder
    lf0 = LF0Encoder(c_h=192, c_out=192, c_out_g=192, num_layer=2, c_in=1)
    lf0.load_state_dict(synth_lf0_weights(prefix=""), strict=True)
    tv = TVEncoder(c_in=80, c_out=192, c_out_g=192, num_layer=6, c_h=128, n_emb=512, commit_w=0.25)
    tv.load_state_dict(synth_tv_weights(prefix=""), strict=True)
    tiv = TIVEncoder(c_in=80, c_out=64, num_layer=6, c_h=128)
    tiv.load_state_dict(synth_tiv_weights(prefix=""), strict=True)
    conv_sty = torch.nn.Conv1d(192, 128, 1, 1)
    cw = synth_conv_sty_weights()
    conv_sty.load_state_dict({"weight": cw["conv_sty.weight"], "bias": cw["conv_sty.bias"]})
    return lf0.cuda().eval(), tv.cuda().eval(), tiv.cuda().eval(), conv_sty.cuda().eval()


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_lf0_encoder_and_fusion_match_reference_fixture(path):
    from dexb200.model import style_fusion
    g = np.load(path)
    B, T, ragged, seed = [int(v) for v in g["meta"]]
    inp = synth_lf0(B, T, seed=seed, ragged=bool(ragged))
    sty = synth_ref_mel(B, T, seed=seed + 100, ragged=bool(ragged))
    lf0, tv, _, conv_sty = make_modules()
    le, ld = lf0(inp["lf0"].cuda(), inp["mask"].cuda())
    assert lf0.cuda_engine().launches == 1 + 2 * 2 + 8
    zb, zd, _ = tv(sty["ref"].cuda(), sty["mask"].cuda())
    se, sd = style_fusion(conv_sty, zb, zd, sty["mask"].cuda(), le, ld, inp["mask"].cuda())
    errs = {k: tensor_rel_err(v.cpu(), torch.from_numpy(g[k])) for k, v in (("lf0_enc", le), ("lf0_dec", ld), ("sty_enc", se), ("sty_dec", sd))}
    print(f"lf0 fixture {os.path.basename(path)}: " + " ".join(f"{k} {e:.2e}" for k, e in errs.items()))
    assert se.shape == (B, 192) and sd.shape == (B, 128, T)
    assert max(errs.values()) < TOL, errs


@pytest.mark.parametrize("B,T,ragged", [(2, 259, True), (1, 1, False), (3, 97, True)])
def test_lf0_encoder_matches_oracle(B, T, ragged):
    inp = synth_lf0(B, T, seed=300 + T, ragged=ragged)
    lf0, _, _, _ = make_modules()
    le, ld = lf0(inp["lf0"].cuda(), inp["mask"].cuda())
    with torch.no_grad():
        le_ref, ld_ref = O.lf0_encoder(synth_lf0_weights(), inp["lf0"], inp["mask"])
    e1, e2 = tensor_rel_err(le.cpu(), le_ref), tensor_rel_err(ld.cpu(), ld_ref)
    print(f"lf0 B={B} T={T}: lf0_enc {e1:.2e} lf0_dec {e2:.2e}")
    assert e1 < TOL and e2 < TOL
    le2, ld2 = lf0(inp["lf0"].cuda(), inp["mask"].cuda())
    assert torch.equal(le2, le) and torch.equal(ld2, ld)


def test_pre_loop_stage_feeds_the_loop():
    """DeXTTS.forward order (tts.py:38-50,71) on the GPU: lf0 / sty / ref -> LF0, TV, TIV encoders -> style fusion -> decoder,
    against the same chain of the CPU oracle."""
    from dexb200.engine import ReverseDiffusion
    from dexb200.model import style_fusion
    cfg = DecoderCfg.make("dex")
    w = synth_decoder_weights(cfg, seed=100, live=True)
    B, T, Ts, steps = 2, 64, 45, 4
    inp = synth_inputs(cfg, B, T, Ts=Ts, seed=8, ragged=True)
    ref = synth_ref_mel(B, Ts, seed=9, ragged=True)                 # synthesize.py: ref = sty = the reference mel
    f0 = synth_lf0(B, Ts, seed=10, ragged=False)
    f0["mask"], f0["lf0"] = ref["mask"], f0["lf0"] * ref["mask"].squeeze(1)
    lf0, tv, tiv, conv_sty = make_modules()
    le, ld = lf0(f0["lf0"].cuda(), f0["mask"].cuda())
    zb, zd, _ = tv(ref["ref"].cuda(), ref["mask"].cuda())
    _, sty = style_fusion(conv_sty, zb, zd, ref["mask"].cuda(), le, ld, f0["mask"].cuda(), want_sty_enc=False)
    _, skips = tiv(ref["ref"].cuda(), ref["mask"].cuda())
    eng = ReverseDiffusion(cfg)
    eng.load_state_dict(w)
    x0 = inp["z"] / 1.5 + inp["mu"]
    y = eng.sample(x0.cuda(), inp["mask"].cuda(), inp["mu"].cuda(), steps,
                   cond=dict(sty=sty, sty_lengths=ref["ref_lengths"].cuda(), ref_skips=skips)).cpu()
    we = dict(synth_lf0_weights())
    we.update(synth_tv_weights())
    we.update(synth_tiv_weights())
    we.update(synth_conv_sty_weights())
    with torch.no_grad():
        le_r, ld_r = O.lf0_encoder(we, f0["lf0"], f0["mask"])
        zb_r, zd_r, _ = O.tv_encoder(we, ref["ref"], ref["mask"])
        _, sty_r = O.style_fusion(we, zb_r, zd_r, le_r, ld_r, ref["mask"], f0["mask"])
        _, skips_r = O.tiv_encoder(we, ref["ref"], ref["mask"])
        y_ref = O.reverse_diffusion(w, O.make_cfg("dex"), inp["z"], inp["mask"], inp["mu"], steps, temperature=1.5,
                                    cond=dict(sty=sty_r, sty_lengths=ref["ref_lengths"], ref_skips=skips_r))
    e_sty = tensor_rel_err(sty.cpu(), sty_r)
    v = per_bin_violation(y, y_ref)
    print(f"pre-loop chain: sty err {e_sty:.2e}, mel per-bin violation {v:.2e}")
    assert e_sty < TOL
    assert v < REL_TOL
    eng.close()
