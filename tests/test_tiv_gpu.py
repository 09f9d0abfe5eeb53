"""GPU parity of the TIV encoder (dexb_tiv_* through the C ABI) against the CPU oracle and the reference fixtures, and the chain
TIV encoder -> reverse diffusion (the encoder's skips consumed by the loop's TIVAdaptor) against the oracle chain."""
import glob
import os

import numpy as np
import pytest
import torch

import dex_oracle as O
from dexb200.manifest import DecoderCfg
from dexb200.synth import synth_decoder_weights, synth_inputs, synth_ref_mel, synth_tiv_weights
from parity import REL_TOL, per_bin_violation, tensor_rel_err

pytestmark = pytest.mark.gpu

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "tiv_*.npz")))
# Tolerance = the path's 1e-3 (BASELINE.json north_star), taken as max |a - b| over the RMS of the reference tensor.  Measured on
# B200: ~3e-4 for the deepest tensors (split-bf16 x3 convolutions contribute ~2e-5 of the RMS each -- tests/test_gemm_gpu.py --
# and every InstanceNorm1D divides low-variance channels, and their error, by a small standard deviation).
TIV_TOL = 1e-3


def make_module():
    from dexb200.model import TIVEncoder
    m = TIVEncoder(c_in=80, c_out=64, num_layer=6, c_h=128)
    m.load_state_dict(synth_tiv_weights(prefix=""), strict=True)
    return m.cuda().eval()


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_tiv_encoder_matches_reference_fixture(path):
    g = np.load(path)
    B, T, ragged, seed = [int(v) for v in g["meta"]]
    inp = synth_ref_mel(B, T, seed=seed, ragged=bool(ragged))
    m = make_module()
    out, skips = m(inp["ref"].unsqueeze(1).cuda(), inp["mask"].cuda())          # (B,1,80,T) like synthesize.py feeds it
    torch.cuda.synchronize()
    assert m.cuda_engine().launches == 3 + 4 * 6 + 2
    errs = [tensor_rel_err(s.cpu(), torch.from_numpy(g[f"skip{i}"])) for i, s in enumerate(skips)]
    errs.append(tensor_rel_err(out.cpu(), torch.from_numpy(g["out"])))
    print(f"tiv fixture {os.path.basename(path)}: max err / rms of skips 0..5, out = " + " ".join(f"{e:.2e}" for e in errs))
    assert all(s.shape == g[f"skip{i}"].shape for i, s in enumerate(skips))
    assert max(errs) < TIV_TOL, errs
    pad = (1.0 - inp["mask"]).cuda()
    assert float((skips[-1] * pad).abs().max()) == 0.0


@pytest.mark.parametrize("B,T,ragged", [(2, 259, True), (1, 2, False), (5, 128, True), (2, 130, False)])
def test_tiv_encoder_matches_oracle(B, T, ragged):
    inp = synth_ref_mel(B, T, seed=100 + T, ragged=ragged)
    m = make_module()
    out, skips = m(inp["ref"].cuda(), inp["mask"].cuda())
    with torch.no_grad():
        out_ref, skips_ref = O.tiv_encoder(synth_tiv_weights(), inp["ref"], inp["mask"])
    assert all(bool(torch.isfinite(s).all()) for s in skips) and bool(torch.isfinite(out).all())
    # T = 2 (the minimum InstanceNorm1D accepts): two-frame statistics amplify rounding noise wherever the two frames of a channel
    # nearly coincide, so only the first skip (taken before any normalisation) is compared there
    n_cmp = 6 if T > 2 else 1
    errs = [tensor_rel_err(s.cpu(), r) for s, r in zip(skips[:n_cmp], skips_ref[:n_cmp])]
    if T > 2:
        errs.append(tensor_rel_err(out.cpu(), out_ref))
    print(f"tiv B={B} T={T}: max err / rms = " + " ".join(f"{e:.2e}" for e in errs))
    assert max(errs) < TIV_TOL, errs
    # a second call with the same shape reuses the plan and reproduces the result bit for bit
    out2, skips2 = m(inp["ref"].cuda(), inp["mask"].cuda())
    assert torch.equal(out2, out) and all(torch.equal(a, b) for a, b in zip(skips2, skips))


def test_tiv_encoder_repacks_changed_weights():
    inp = synth_ref_mel(2, 64, seed=3, ragged=True)
    m = make_module()
    _, s1 = m(inp["ref"].cuda(), inp["mask"].cuda())
    with torch.no_grad():
        m.in_conv.bn.running_var.mul_(4.0)
    _, s2 = m(inp["ref"].cuda(), inp["mask"].cuda())
    w = synth_tiv_weights()
    w["tiv_encoder.in_conv.bn.running_var"] = w["tiv_encoder.in_conv.bn.running_var"] * 4.0
    with torch.no_grad():
        _, s_ref = O.tiv_encoder(w, inp["ref"], inp["mask"])
    assert not torch.equal(s1[0], s2[0])
    assert tensor_rel_err(s2[0].cpu(), s_ref[0]) < TIV_TOL


def test_tiv_encoder_feeds_the_loop():
    """DeXTTS.forward order (tts.py:50,71): ref -> tiv_encoder -> ref_skips -> decoder, all on the GPU, vs the oracle chain."""
    from dexb200.engine import ReverseDiffusion
    cfg = DecoderCfg.make("dex")
    w = synth_decoder_weights(cfg, seed=100, live=True)
    B, T, Ts, Tr, steps = 2, 64, 31, 45, 4
    inp = synth_inputs(cfg, B, T, Ts=Ts, seed=8, ragged=True)
    r = synth_ref_mel(B, Tr, seed=9, ragged=True)
    m = make_module()
    _, skips = m(r["ref"].cuda(), r["mask"].cuda())
    eng = ReverseDiffusion(cfg)
    eng.load_state_dict(w)
    x0 = inp["z"] / 1.5 + inp["mu"]
    y = eng.sample(x0.cuda(), inp["mask"].cuda(), inp["mu"].cuda(), steps,
                   cond=dict(sty=inp["sty"].cuda(), sty_lengths=inp["sty_lengths"].cuda(), ref_skips=skips)).cpu()
    with torch.no_grad():
        _, skips_ref = O.tiv_encoder(synth_tiv_weights(), r["ref"], r["mask"])
        y_ref = O.reverse_diffusion(w, O.make_cfg("dex"), inp["z"], inp["mask"], inp["mu"], steps, temperature=1.5,
                                    cond=dict(sty=inp["sty"], sty_lengths=inp["sty_lengths"], ref_skips=skips_ref))
    assert per_bin_violation(y, y_ref) < REL_TOL
    eng.close()
