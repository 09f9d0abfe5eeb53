"""The top-level drop-ins ``dexb200.model.DeXTTS`` / ``GeDEXTTS`` end to end on the GPU, called the way DEX-TTS/synthesize.py:105 calls
the reference model, against the fixtures of the unmodified reference ``forward`` (tests/golden/tts_*.npz).

What is asserted: the text side (enc_out, the hard alignment) is held to the path tolerance / to equality of the durations; the decoder
is compared with the CPU oracle of the loop fed with the GPU's OWN conditioning (mu_y, sty, ref_skips captured at ``model.decoder``),
AND -- whenever the alignment equals the fixture's -- the final ``dec_out`` with the unmodified reference's ``dec_out`` at the path
tolerance (measured on B200, profiles/r02_pytest_gpu_c.log: 7.9e-5 / 1.9e-4 / 1.5e-4 of the RMS; on these random weights two sampler
steps amplify a 1e-6 perturbation of ``sty`` a hundredfold, tests/test_tts_oracle.py, which is still inside the bound)."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

import dex_oracle as O
from dexb200.synth import seeded_noise
from parity import REL_TOL, per_bin_violation, tensor_rel_err
from test_tts_module_cpu import build

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from make_golden_tts import synth_tts_inputs  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.run_last]

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "tts_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_model_forward_on_gpu(path, monkeypatch):
    g = np.load(path)
    variant = str(g["variant"])
    B, Tx, Ts, steps, ragged, seed = [int(v) for v in g["meta"]]
    temperature, length_scale = [float(v) for v in g["scale"]]
    inp = {k: v.cuda() for k, v in synth_tts_inputs(variant, B, Tx, Ts, seed, bool(ragged)).items() if torch.is_tensor(v)}
    model, w = build(variant)
    model = model.cuda().eval()
    cap = {}
    model.decoder.register_forward_pre_hook(lambda m, args, kwargs: cap.update(args=args, kwargs=kwargs), with_kwargs=True)
    noise = seeded_noise(seed + 3)
    drawn = []
    real_randn = torch.randn

    def fake_randn(*shape, **kw):                          # Diffusion.forward's on-device draw (diffusion.py:256) -> the fixture's noise
        shp = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
        z = noise(shp)
        drawn.append(z)
        return z.to(kw.get("device", "cpu"))
    monkeypatch.setattr(torch, "randn", fake_randn)
    if variant == "dex":
        enc_out, dec_out, attn = model(inp["x"], inp["x_lengths"], inp["ref"], inp["ref_lengths"], inp["ref"], inp["ref_lengths"],
                                       inp["lf0"], inp["lf0_lengths"], spk=None, n_timesteps=steps, temperature=temperature,
                                       length_scale=length_scale)
    else:
        enc_out, dec_out, attn = model(inp["x"], inp["x_lengths"], n_timesteps=steps, temperature=temperature, spk=None,
                                       length_scale=length_scale)
    torch.cuda.synchronize()
    monkeypatch.setattr(torch, "randn", real_randn)
    assert len(drawn) == 1 and torch.isfinite(dec_out).all() and enc_out.shape == dec_out.shape

    # decoder parity on the GPU's own conditioning
    a = cap["args"]
    mu_y, y_mask = a[0].cpu(), a[1].cpu()
    cond = dict(sty=a[5].cpu(), sty_lengths=a[6].cpu(), ref_skips=[r.cpu() for r in a[3]]) if variant == "dex" else None
    with torch.no_grad():
        y = O.reverse_diffusion(w, O.make_cfg(variant), drawn[0], y_mask, mu_y, steps, temperature, cond)
    v_dec = per_bin_violation(dec_out.cpu(), y[:, :, :dec_out.shape[-1]])
    print(f"{os.path.basename(path)}: decoder vs oracle on the same conditioning {v_dec:.2e}")
    assert v_dec < REL_TOL

    # text side against the reference fixture
    shape = tuple(int(n) for n in g["attn_shape"])
    attn_ref = np.unpackbits(g["attn"], axis=-1, count=shape[-1]).astype(np.float32).reshape(shape)
    if tuple(attn.shape) == shape and np.array_equal(attn.cpu().numpy(), attn_ref):
        v_enc = per_bin_violation(enc_out.cpu(), torch.from_numpy(g["enc_out"]))
        e_dec = tensor_rel_err(dec_out.cpu(), torch.from_numpy(g["dec_out"]))
        print(f"{os.path.basename(path)}: alignment identical, enc_out {v_enc:.2e}, dec_out vs the reference fixture {e_dec:.2e}")
        assert v_enc < REL_TOL and e_dec < REL_TOL
    else:                                                  # a duration flipped: only legitimate within noise of a rounding boundary
        dur = attn.squeeze(1).sum(-1).cpu().numpy()
        dur_ref = attn_ref[:, 0].sum(-1)
        n_flip = int((dur[:, :dur_ref.shape[1]] != dur_ref[:, :dur.shape[1]]).sum())
        print(f"{os.path.basename(path)}: {n_flip} duration(s) differ from the reference fixture")
        assert n_flip <= 1


def test_multi_speaker_gedex_model_forward_on_gpu(monkeypatch):
    """GeDEX-TTS with n_spks > 1 (GeDEX-TTS/config/VCTK/base.yaml: 108 speakers) end to end: spk ids -> spk_emb -> text encoder (speaker
    channel behind the prenet, 256 wide) -> alignment -> decoder (spk_mlp(spk) as third input channel), G/model/tts.py:27-56.  Checked
    against the oracle chain on the same seeded weights (each oracle stage is pinned bit-exactly against the unmodified reference on
    the n_spks > 1 fixtures text_gedex_spk_b2r / gedex_spk_b2r)."""
    import text_oracle as TO
    from dexb200.manifest import DecoderCfg
    from dexb200.model import GeDEXTTS
    from dexb200.synth import reference_state_dict, synth_decoder_weights, synth_text, synth_text_weights
    from test_tts_module_cpu import CFG
    n_spks = 4
    cfg = dict(CFG["gedex"], n_spks=n_spks)
    dcfg = DecoderCfg.make("gedex", n_spks=n_spks)
    w = dict(synth_decoder_weights(dcfg, seed=100, live=True))
    w.update(synth_text_weights(seed=100, adaln=False, spk_emb_dim=64))
    emb = torch.randn(n_spks, 64, generator=torch.Generator().manual_seed(9)) * 0.5
    sd = reference_state_dict(w)
    sd["spk_emb.weight"] = emb
    model = GeDEXTTS(cfg)
    assert set(model.state_dict().keys()) == set(sd.keys())
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    B, Tx, steps, temperature = 2, 40, 3, 1.5
    inp = synth_text(B, Tx, seed=4242, ragged=True)
    spk_id = torch.tensor([3, 1])
    noise = seeded_noise(77)
    drawn = []
    real_randn = torch.randn

    def fake_randn(*shape, **kw):
        shp = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
        z = noise(shp)
        drawn.append(z)
        return z.to(kw.get("device", "cpu"))
    monkeypatch.setattr(torch, "randn", fake_randn)
    enc_out, dec_out, attn = model(inp["x"].cuda(), inp["x_lengths"].cuda(), n_timesteps=steps, temperature=temperature, spk=spk_id.cuda())
    torch.cuda.synchronize()
    monkeypatch.setattr(torch, "randn", real_randn)
    assert len(drawn) == 1
    with torch.no_grad():
        spk = emb[spk_id]
        mu_x, logw, x_mask = TO.text_encoder(w, inp["x"], inp["x_lengths"], None, spk=spk)
        mu_y, y_mask, r_attn, _, y_max = O.align_durations(logw, x_mask, mu_x, 1.0)
        if tuple(attn.shape) == tuple(r_attn[:, :, :y_max].shape) and torch.equal(attn.cpu(), r_attn[:, :, :y_max]):
            y = O.reverse_diffusion(w, O.make_cfg("gedex", n_spks=n_spks), drawn[0], y_mask, mu_y, steps, temperature, dict(spk=spk))
            v_enc = per_bin_violation(enc_out.cpu(), mu_y[:, :, :y_max])
            v_dec = per_bin_violation(dec_out.cpu(), y[:, :, :y_max])
            print(f"multi-speaker GeDEX-TTS: enc_out {v_enc:.2e}, dec_out {v_dec:.2e} (vs the oracle chain)")
            assert v_enc < REL_TOL and v_dec < REL_TOL
        else:
            dur, r_dur = attn.squeeze(1).sum(-1).cpu(), r_attn.squeeze(1).sum(-1)
            n_flip = int((dur != r_dur[:, :dur.shape[1]]).sum())
            print(f"multi-speaker GeDEX-TTS: {n_flip} duration(s) differ from the oracle chain")
            assert n_flip <= 1
