"""GPU parity of the text encoder (dexb_text_* through the C ABI, drop-in ``model.TextEncoder`` / ``GeTextEncoder``) against the
fixtures of the unmodified reference modules (tests/golden/text_*.npz) and the CPU oracle, and chained into the duration / alignment
glue the way DeXTTS.forward orders it (tts.py:51-68).  The fixture comparisons below are what tools/text_check.py ran on a B200
(profiles/r01_text_check_v32.log: mu 8e-5 ... 1.2e-4, logw 6e-5, residual stream after the last layer <= 2e-4 of the RMS)."""
import glob
import os

import numpy as np
import pytest
import torch

import dex_oracle as O
import text_oracle as TO
from dexb200.synth import synth_text, synth_text_weights
from parity import tensor_rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.run_last]

GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "text_*.npz"))
              if "_spk_" not in os.path.basename(p))       # the n_spks > 1 fixture has its own test below (speaker input)
GOLD_SPK = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "text_*_spk_*.npz")))
TOL = 1e-3             # the path tolerance (north_star), as max |a - b| / RMS(reference tensor)
KW = dict(n_vocab=149, n_feats=80, n_channels=192, filter_channels=1024, filter_channels_dp=256, n_heads=2, n_layers=8, kernel_size=3,
          p_dropout=0.1, use_softmax=True, use_decay=False, window_size=4)        # DEX-TTS/config/VCTK/base.yaml:51-61


def make_encoder(dex=True):
    from dexb200.model import GeTextEncoder, TextEncoder
    enc = (TextEncoder if dex else GeTextEncoder)(**KW)
    enc.load_state_dict(synth_text_weights(prefix="", adaln=dex), strict=True)
    return enc.cuda().eval()


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_text_encoder_matches_reference_fixture(path):
    g = np.load(path)
    B, Tx, ragged, seed, dex = [int(v) for v in g["meta"][:5]]
    inp = synth_text(B, Tx, seed=seed, ragged=bool(ragged))
    enc = make_encoder(bool(dex))
    x, xl, sty = inp["x"].cuda(), inp["x_lengths"].cuda(), inp["sty"].cuda()
    mu, logw, x_mask = enc(x, xl, sty) if dex else enc(x, xl)
    eng = enc.cuda_engine()
    assert eng.launches == (1 if dex else 0) + 1 + 6 + 3 + 8 * 9 + 7          # 9 launches per layer: q|k|v|g and fc1|gate are one GEMM each
    assert mu.shape == (B, 80, Tx) and logw.shape == (B, 1, Tx) and np.array_equal(x_mask.cpu().numpy(), g["x_mask"])
    s = sty if dex else None
    errs = {"mu": tensor_rel_err(mu.cpu(), torch.from_numpy(g["mu"])), "logw": tensor_rel_err(logw.cpu(), torch.from_numpy(g["logw"])),
            "prenet": tensor_rel_err(eng.forward_stream(x, x_mask, s, 0).cpu(), torch.from_numpy(g["prenet"]).transpose(1, 2)),
            "layer0": tensor_rel_err(eng.forward_stream(x, x_mask, s, 1).cpu(), torch.from_numpy(g["layer0"])),
            "layer7": tensor_rel_err(eng.forward_stream(x, x_mask, s, 8).cpu(), torch.from_numpy(g["layer7"]))}
    print(f"text fixture {os.path.basename(path)}: " + " ".join(f"{k} {e:.2e}" for k, e in errs.items()))
    assert max(errs.values()) < TOL, errs
    pad = (1 - x_mask).cpu()
    assert float((mu.cpu() * pad).abs().max()) == 0.0 and float((logw.cpu() * pad).abs().max()) == 0.0
    mu2, logw2, _ = enc(x, xl, sty) if dex else enc(x, xl)          # the layer limit of forward_stream is restored; reproducible
    assert torch.equal(mu, mu2) and torch.equal(logw, logw2)


@pytest.mark.parametrize("path", GOLD_SPK, ids=[os.path.basename(p)[:-4] for p in GOLD_SPK])
def test_multi_speaker_text_encoder_matches_reference_fixture(path):
    """n_spks > 1 (GeDEX-TTS/config/VCTK: 108 speakers): the speaker embedding joins behind the prenet, the RetNet / proj_m / duration
    predictor run 256 wide with head dim 128 (GeDEX-TTS/model/text_encoder.py:119-127,141-142).  Fixture = the unmodified reference."""
    from dexb200.model import GeTextEncoder
    g = np.load(path)
    B, Tx, ragged, seed, dex, n_spks = [int(v) for v in g["meta"][:6]]
    assert not dex and n_spks > 1
    inp = synth_text(B, Tx, seed=seed, ragged=bool(ragged))
    enc = GeTextEncoder(**KW, spk_emb_dim=64, n_spks=n_spks)
    sd = synth_text_weights(prefix="", adaln=False, spk_emb_dim=64)
    assert list(sd.keys()) == list(g["keys"]) == list(enc.state_dict().keys())
    enc.load_state_dict(sd, strict=True)
    enc = enc.cuda().eval()
    spk = torch.randn(B, 64, generator=torch.Generator().manual_seed(seed + 7))          # oracle/make_golden_text.py
    x, xl = inp["x"].cuda(), inp["x_lengths"].cuda()
    mu, logw, x_mask = enc(x, xl, spk=spk.cuda())
    eng = enc.cuda_engine()
    assert eng.launches == 1 + 6 + 3 + 1 + 8 * 9 + 7
    errs = {"mu": tensor_rel_err(mu.cpu(), torch.from_numpy(g["mu"])), "logw": tensor_rel_err(logw.cpu(), torch.from_numpy(g["logw"])),
            "layer0": tensor_rel_err(eng.forward_stream(x, x_mask, None, 1, spk=spk.cuda()).cpu(), torch.from_numpy(g["layer0"])),
            "layer7": tensor_rel_err(eng.forward_stream(x, x_mask, None, 8, spk=spk.cuda()).cpu(), torch.from_numpy(g["layer7"]))}
    pre = eng.forward_stream(x, x_mask, None, 0, spk=spk.cuda()).cpu()                      # (B, Tx, 192 + 64)
    errs["prenet"] = tensor_rel_err(pre[:, :, :192], torch.from_numpy(g["prenet"]).transpose(1, 2))
    assert torch.equal(pre[:, :, 192:], spk[:, None, :].expand(B, Tx, 64))                  # repeated over ALL positions, unmasked
    print(f"text fixture {os.path.basename(path)}: " + " ".join(f"{k} {e:.2e}" for k, e in errs.items()))
    assert max(errs.values()) < TOL, errs
    with pytest.raises(RuntimeError):
        enc(x, xl)                                           # the speaker embedding is required for n_spks > 1


@pytest.mark.parametrize("B,Tx,ragged", [(8, 128, True), (2, 512, True), (1, 1, False)])
def test_text_encoder_matches_oracle(B, Tx, ragged):
    """Phoneme lengths of BASELINE.json's configs (128: C2 / C3, 512: C5: four 128-token tiles per utterance) and a single token."""
    inp = synth_text(B, Tx, seed=100 + B + Tx, ragged=ragged)
    enc = make_encoder(True)
    mu, logw, x_mask = enc(inp["x"].cuda(), inp["x_lengths"].cuda(), inp["sty"].cuda())
    with torch.no_grad():
        r_mu, r_logw, r_mask = TO.text_encoder(synth_text_weights(), inp["x"], inp["x_lengths"], inp["sty"])
    e_mu, e_lw = tensor_rel_err(mu.cpu(), r_mu), tensor_rel_err(logw.cpu(), r_logw)
    print(f"text B={B} Tx={Tx}: mu {e_mu:.2e} logw {e_lw:.2e}")
    assert torch.equal(x_mask.cpu(), r_mask) and e_mu < TOL and e_lw < TOL


def test_text_encoder_feeds_the_alignment_glue():
    """DeXTTS.forward order (tts.py:51-68) on the GPU: ids, sty -> TextEncoder -> align_durations -> mu_y, vs the oracle chain.
    ceil(exp(logw)) may legitimately differ where the oracle's own duration sits within the arithmetic noise of an integer."""
    from dexb200.model import align_durations
    inp = synth_text(4, 64, seed=321, ragged=True)
    enc = make_encoder(True)
    mu_x, logw, x_mask = enc(inp["x"].cuda(), inp["x_lengths"].cuda(), inp["sty"].cuda())
    mu_y, y_mask, attn, y_lengths, y_max = align_durations(logw, x_mask, mu_x)
    with torch.no_grad():
        r_mu_x, r_logw, r_mask = TO.text_encoder(synth_text_weights(), inp["x"], inp["x_lengths"], inp["sty"])
        r_mu_y, r_y_mask, r_attn, r_len, r_max = O.align_durations(r_logw, r_mask, r_mu_x)
    dur = attn.squeeze(1).sum(-1).cpu()
    r_w = (torch.exp(r_logw) * r_mask).squeeze(1)
    r_dur = torch.ceil(r_w)
    near_integer = (r_w - torch.round(r_w)).abs() < 2e-3 * r_w.clamp_min(1.0)
    assert bool(((dur == r_dur) | near_integer).all()), "durations differ away from a rounding boundary"
    if torch.equal(dur, r_dur):
        assert y_max == r_max and torch.equal(y_lengths.cpu(), r_len) and torch.equal(attn.cpu(), r_attn)
        err = tensor_rel_err(mu_y.cpu(), r_mu_y)
        print(f"text -> align chain: mu_y {err:.2e}, y_lengths {y_lengths.tolist()}")
        assert err < TOL
