"""Whole-network parity AT THE BENCHMARKED SHAPES (BASELINE.json configs C2-C5): one preconditioned network call and a short
trajectory of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

At these sizes different code runs than in the small fixtures: four halo tiles per image row, the attention tail split, split-KV
linear-attention contexts over 40 960 keys, the 640-CTA TV adaptor, the positional conv at Wq = 129 / 501, 32-bit index math of the
element-wise kernels.  Utterances are independent, so for the large batches the oracle runs on a few samples of the
batch, each as a batch of one at the same padded length.

Tolerance: BASELINE.json north_star, 1e-3 relative fp32 per mel bin (tests/parity.py).
Reference: Diffusion.forward(infer=True) -> ablation_sampler -> EDMPrecond -> DiffusionDenoiser
(DEX-TTS/model/diffusion.py:190-259, DEX-TTS/model/edm.py:88-98,183-209)."""
import os

import pytest
import torch

import dex_oracle as O
from dexb200.manifest import DecoderCfg
from dexb200.synth import synth_decoder_weights, synth_inputs
from parity import REL_TOL, per_bin_violation
from test_decoder_gpu import get_engine, to_cuda

pytestmark = pytest.mark.gpu


def _threads():
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))


def _cond(variant, inp):
    return dict(sty=inp["sty"], sty_lengths=inp["sty_lengths"], ref_skips=inp["ref_skips"]) if variant == "dex" else None


def _slice(inp, cond, i):
    s = slice(i, i + 1)
    c = None if cond is None else dict(sty=cond["sty"][s], sty_lengths=cond["sty_lengths"][s], ref_skips=[r[s] for r in cond["ref_skips"]])
    return inp["z"][s], inp["mask"][s], inp["mu"][s], c


def _check_call(variant, B, T, Ts, n_steps, step, samples, ragged=False, seed=1234):
    """D(x; sigma_step) of the CUDA path for the whole batch vs the oracle on `samples`."""
    _threads()
    cfg = DecoderCfg.make(variant)
    w = synth_decoder_weights(cfg, seed=100, live=True)
    inp = synth_inputs(cfg, B, T, Ts=max(Ts, 1), seed=seed, ragged=ragged)
    cond = _cond(variant, inp)
    ts = O.sigma_schedule(n_steps)
    # a state typical of that noise level: data + sigma * noise (step 0: exactly the sampler's initial state)
    x = (inp["z"] / 1.5 + inp["mu"]) * ts[0] if step == 0 else inp["mu"] + ts[step] * inp["z"]
    eng = get_engine(variant, True, 0)
    out = eng.denoise_once(x.cuda(), inp["mask"].cuda(), inp["mu"].cuda(), n_steps, step, cond=to_cuda(cond)).cpu()
    assert torch.isfinite(out).all()
    ocfg = O.make_cfg(variant)
    worst = 0.0
    for i in samples:
        z, m, mu, c = _slice(inp, cond, i)
        with torch.no_grad():
            ref = O.edm_precond(w, ocfg, x[i:i + 1], ts[step], m, mu, cond=c)
        worst = max(worst, per_bin_violation(out[i:i + 1], ref))
    print(f"{variant} B={B} T={T} step {step}/{n_steps} samples {list(samples)}: per-bin violation {worst:.3e} (tol {REL_TOL:g})")
    assert worst < REL_TOL
    assert eng.simt_fallbacks == 0


def _check_traj(variant, B, T, Ts, n_steps, samples, ragged=False, seed=1234, cond_override=None):
    _threads()
    cfg = DecoderCfg.make(variant)
    w = synth_decoder_weights(cfg, seed=100, live=True)
    inp = synth_inputs(cfg, B, T, Ts=max(Ts, 1), seed=seed, ragged=ragged)
    cond = _cond(variant, inp) if cond_override is None else cond_override
    eng = get_engine(variant, True, 0)
    x0 = inp["z"] / 1.5 + inp["mu"]
    y = eng.sample(x0.cuda(), inp["mask"].cuda(), inp["mu"].cuda(), n_steps, cond=to_cuda(cond)).cpu()
    assert torch.isfinite(y).all()
    ocfg = O.make_cfg(variant)
    worst = 0.0
    for i in samples:
        z, m, mu, c = _slice(inp, cond, i)
        with torch.no_grad():
            ref = O.reverse_diffusion(w, ocfg, z, m, mu, n_steps, temperature=1.5, cond=c)
        worst = max(worst, per_bin_violation(y[i:i + 1], ref))
    print(f"{variant} B={B} T={T} {n_steps}-step trajectory samples {list(samples)}: per-bin violation {worst:.3e} (tol {REL_TOL:g})")
    assert worst < REL_TOL
    return y


# ---- C2: DEX-TTS B=8, T=512, Ts=259, 50-step schedule (the configuration BENCH is quoted on)
@pytest.mark.parametrize("step", [0, 25])
def test_c2_network_call(step):
    _check_call("dex", 8, 512, 259, 50, step, range(8) if step == 0 else (0, 5))


def test_c2_three_step_trajectory():
    _check_traj("dex", 8, 512, 259, 3, range(8))


def test_c2_ragged_two_step_trajectory():
    _check_traj("dex", 8, 512, 259, 2, (0, 3, 7), ragged=True, seed=4321)


# ---- C4: GeDEX-TTS B=32 per GPU, T=512
def test_c4_network_call_and_trajectory():
    _check_call("gedex", 32, 512, 0, 50, 0, (0, 13, 31))
    _check_traj("gedex", 32, 512, 0, 2, (0, 13, 31))


# ---- C5: DEX-TTS long-form, B=8 per GPU, T=2000 (10 020 DiT tokens)
def test_c5_network_call_and_trajectory():
    _check_call("dex", 8, 2000, 259, 50, 0, (5,))
    _check_traj("dex", 8, 2000, 259, 2, (2,))


# ---- C3: DEX-TTS B=32, conditioning from the CUDA front-end (STFT -> mel -> TIV / TV / LF0 encoders -> fusion) on 3 s of audio
def test_c3_batch32_conditioning_from_stft():
    from dexb200.engine import stft_mel
    from dexb200.model import LF0Encoder, TIVEncoder, TVEncoder, style_fusion
    from dexb200.synth import synth_conv_sty_weights, synth_lf0, synth_lf0_weights, synth_tiv_weights, synth_tv_weights
    import stft_oracle as SO
    B, S = 32, 66150
    g = torch.Generator().manual_seed(99)
    audio = (torch.rand(B, S, generator=g) - 0.5).cuda()
    win = torch.hann_window(1024, periodic=True).cuda()
    fb = torch.from_numpy(SO.mel_filterbank()).float().cuda()
    mel = stft_mel(audio, win, fb)
    n_ref = S // 256 + 1
    assert mel.shape == (B, 80, n_ref) and n_ref == 259
    tiv = TIVEncoder(c_in=80, c_out=64, num_layer=6, c_h=128)
    tiv.load_state_dict(synth_tiv_weights(prefix=""), strict=True)
    tv = TVEncoder(c_in=80, c_out=192, c_out_g=192, num_layer=6, c_h=128, n_emb=512, commit_w=0.25)
    tv.load_state_dict(synth_tv_weights(prefix=""), strict=True)
    lf0e = LF0Encoder(c_h=192, c_out=192, c_out_g=192, num_layer=2, c_in=1)
    lf0e.load_state_dict(synth_lf0_weights(prefix=""), strict=True)
    conv_sty = torch.nn.Conv1d(192, 128, 1, 1)
    cw = synth_conv_sty_weights()
    conv_sty.load_state_dict({"weight": cw["conv_sty.weight"], "bias": cw["conv_sty.bias"]})
    tiv, tv, lf0e, conv_sty = tiv.cuda().eval(), tv.cuda().eval(), lf0e.cuda().eval(), conv_sty.cuda().eval()
    ref_mask = torch.ones(B, 1, n_ref, device="cuda")
    lf0 = synth_lf0(B, n_ref, seed=77)["lf0"].cuda()
    with torch.no_grad():
        _, skips = tiv(mel, ref_mask)
        z_before, z_dec, _ = tv(mel, ref_mask)
        lf0_enc, lf0_dec = lf0e(lf0, ref_mask)
        _, sty = style_fusion(conv_sty, z_before, z_dec, ref_mask, lf0_enc, lf0_dec, ref_mask, want_sty_enc=False)
    cond = dict(sty=sty.cpu(), sty_lengths=torch.full((B,), n_ref, dtype=torch.long), ref_skips=[s.cpu() for s in skips])
    _check_traj("dex", B, 512, n_ref, 2, (0, 21), cond_override=cond)
