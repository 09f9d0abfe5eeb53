"""Host logic of the top-level drop-ins ``dexb200.model.DeXTTS`` / ``GeDEXTTS`` without a GPU:
 * their ``state_dict`` is the unmodified reference model's, key for key and shape for shape (key lists recorded in the fixtures);
 * ``forward``'s own glue (masks, call order, argument passing, the slices of tts.py:69-74) is exercised with every CUDA stage
   replaced by its oracle counterpart and compared with the outputs of the reference models' ``forward`` (tests/golden/tts_*.npz).
The stages themselves are covered by the ``-m gpu`` tests; nothing here is a product fallback -- the fakes live in this file only."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

import dex_oracle as O
import text_oracle as TO
from dexb200.synth import reference_state_dict, seeded_noise, synth_tts_weights
from parity import REL_TOL, per_bin_violation

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from make_golden_tts import synth_tts_inputs  # noqa: E402

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "tts_*.npz")))

# the `model:` blocks of DEX-TTS/config/VCTK/base.yaml:23-78 and GeDEX-TTS/config/LJSpeech/base.yaml:23-63 (+ n_vocab, set by the scripts)
ENC = dict(n_channels=192, filter_channels=1024, filter_channels_dp=256, n_layers=8, kernel_size=3, p_dropout=0.1, n_heads=2,
           window_size=4, use_softmax=True, use_decay=False)
DEC = dict(dim=64, pe_scale=1000, dim_mults=[1, 2], model_type="dit", precond="edm", loss_type="base")
CFG = {
    "dex": dict(add_blank=True, n_feats=80, n_spks=0, spk_emb_dim=64, n_vocab=149,
                tv_encoder=dict(c_in=80, num_layer=6, c_h=128, c_out=192, c_out_g=192, commit_w=0.25, n_emb=512),
                lf0_encoder=dict(c_in=1, c_h=192, c_out=192, c_out_g=192, num_layer=2),
                tiv_encoder=dict(c_in=80, num_layer=6, c_h=128, c_out=64), encoder=ENC, decoder=DEC,
                dit=dict(in_channels=128, patch_size=3, stride_size=2, overlap=True, hidden_size=256, depth=4, num_heads=2, mlp_ratio=2,
                         out_channels=128, conv_pos=16, conv_pos_groups=8, use_decoder=False, mask_type="time_random")),
    "gedex": dict(add_blank=True, n_feats=80, n_spks=1, spk_emb_dim=64, n_vocab=149, encoder=ENC, decoder=DEC,
                  dit=dict(in_channels=128, patch_size=7, stride_size=4, overlap=True, hidden_size=256, depth=4, num_heads=2, mlp_ratio=2,
                           out_channels=128, conv_pos=16, conv_pos_groups=8, use_decoder=False, mask_type="time_random")),
}


def build(variant):
    from dexb200.model import DeXTTS, GeDEXTTS
    model = (DeXTTS if variant == "dex" else GeDEXTTS)(CFG[variant])
    w = synth_tts_weights(variant)
    model.load_state_dict(reference_state_dict(w), strict=True)
    return model.eval(), w


@pytest.mark.parametrize("variant", ["dex", "gedex"])
def test_state_dict_is_the_reference_models(variant):
    g = np.load([p for p in GOLD if f"tts_{variant}_" in p][0])
    model, _ = build(variant)
    sd = model.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    assert [list(v.shape) for v in sd.values()] == [[int(n) for n in s.split(",") if n] for s in g["shapes"]]


def with_oracle_stages(monkeypatch, model, w, variant, noise):
    """Every CUDA stage of the model -> the oracle function for the same reference code."""
    import dexb200.model.tts as T
    monkeypatch.setattr(T, "align_durations", lambda logw, x_mask, mu_x, length_scale=1.0: O.align_durations(logw, x_mask, mu_x, length_scale))
    cfg = O.make_cfg(variant)
    if variant == "dex":
        monkeypatch.setattr(model.lf0_encoder, "forward", lambda lf0, mask: O.lf0_encoder(w, lf0, mask))
        monkeypatch.setattr(model.tv_encoder, "forward", lambda sty, mask: O.tv_encoder(w, sty, mask))
        monkeypatch.setattr(model.tiv_encoder, "forward", lambda ref, mask: O.tiv_encoder(w, ref, mask))
        monkeypatch.setattr(T, "style_fusion", lambda conv_sty, se, sd, sm, le, ld, lm: O.style_fusion(w, se, sd, le, ld, sm, lm))
        monkeypatch.setattr(model.encoder, "forward", lambda x, xl, sty, spk=None: TO.text_encoder(w, x, xl, sty))
        monkeypatch.setattr(model.decoder, "forward",
                            lambda x, mask, mu, ref, ref_lengths, sty, sty_lengths, n_timesteps=1, spk=None, infer=False, temperature=1.0:
                            O.reverse_diffusion(w, cfg, noise(tuple(mu.shape)), mask, mu, n_timesteps, temperature,
                                                dict(sty=sty, sty_lengths=sty_lengths, ref_skips=ref)))
    else:
        monkeypatch.setattr(model.encoder, "forward", lambda x, xl, spk=None: TO.text_encoder(w, x, xl, None))
        monkeypatch.setattr(model.decoder, "forward", lambda x, mask, mu, n_timesteps=1, spk=None, infer=False, temperature=1.0:
                            O.reverse_diffusion(w, cfg, noise(tuple(mu.shape)), mask, mu, n_timesteps, temperature))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_forward_glue_matches_reference_forward(path, monkeypatch):
    g = np.load(path)
    variant = str(g["variant"])
    B, Tx, Ts, steps, ragged, seed = [int(v) for v in g["meta"]]
    temperature, length_scale = [float(v) for v in g["scale"]]
    inp = synth_tts_inputs(variant, B, Tx, Ts, seed, bool(ragged))
    model, w = build(variant)
    with_oracle_stages(monkeypatch, model, w, variant, seeded_noise(seed + 3))
    if variant == "dex":                                   # the call of DEX-TTS/synthesize.py:105 (ref = sty = the reference mel)
        enc_out, dec_out, attn = model(inp["x"], inp["x_lengths"], inp["ref"], inp["ref_lengths"], inp["ref"], inp["ref_lengths"],
                                       inp["lf0"], inp["lf0_lengths"], spk=None, n_timesteps=steps, temperature=temperature,
                                       length_scale=length_scale)
    else:
        enc_out, dec_out, attn = model(inp["x"], inp["x_lengths"], n_timesteps=steps, temperature=temperature, spk=None,
                                       length_scale=length_scale)
    shape = tuple(int(v) for v in g["attn_shape"])
    attn_ref = np.unpackbits(g["attn"], axis=-1, count=shape[-1]).astype(np.float32).reshape(shape)
    assert tuple(attn.shape) == shape and np.array_equal(attn.numpy(), attn_ref)
    assert enc_out.shape == g["enc_out"].shape and dec_out.shape == g["dec_out"].shape
    assert per_bin_violation(enc_out, torch.from_numpy(g["enc_out"])) < 2e-5
    assert per_bin_violation(dec_out, torch.from_numpy(g["dec_out"])) < REL_TOL        # see tests/test_tts_oracle.py for the bound


REF_ROOT = os.environ.get("DEX_REFERENCE_ROOT", "/root/reference")
needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REF_ROOT, "DEX-TTS", "model")),
                                     reason="the reference checkout is only present in the build container")


def _train_batch(variant, B=2, Tx=17, Ty=44, Ts=23):
    inp = synth_tts_inputs(variant, B, Tx, Ts, 5, True)
    g = torch.Generator().manual_seed(9)
    y = torch.randn(B, 80, Ty, generator=g)
    y_lengths = torch.tensor([Ty, Ty - 8])
    return inp, y, y_lengths


@needs_reference
@pytest.mark.parametrize("variant", ["dex", "gedex"])
def test_compute_loss_delegates_to_the_reference_on_our_parameters(variant, monkeypatch):
    """``compute_loss`` (tts.py:76-153 / GeDEX tts.py:58-121) keeps its signature and is computed by the reference's own modules running
    on the drop-in's parameters: same value as the unmodified reference model with the same weights and RNG state, and gradients land
    on the drop-in's ``nn.Parameter``s.  (MAS runs on the CUDA kernel in production; on this CPU box the oracle stands in.)"""
    import ref_loader
    import mas_oracle
    import dexb200.model.monotonic_align as M
    import dexb200.model.reference_twin as RT
    ref_model, _, cfg = ref_loader.build_reference_tts(variant)                       # installs the timm / transformers shims as well
    mas = lambda value, mask: torch.from_numpy(mas_oracle.maximum_path(value.detach().numpy(), mask.detach().numpy())).to(value.dtype)
    monkeypatch.setattr(M, "maximum_path", mas)
    sys.modules["model.monotonic_align"].maximum_path = mas                          # the stub ref_loader registered for the reference model
    RT.set_reference_dir(os.path.join(REF_ROOT, "DEX-TTS" if variant == "dex" else "GeDEX-TTS"))
    model, w = build(variant)
    ref_model.load_state_dict(reference_state_dict(w), strict=True)
    model.train(); ref_model.train()
    model._training_twin()                       # build the twin now: constructing it draws from the global RNG (parameter init)
    inp, y, y_lengths = _train_batch(variant)
    if variant == "dex":
        args = (inp["x"], inp["x_lengths"], y, y_lengths, inp["ref"], inp["ref_lengths"], inp["ref"], inp["ref_lengths"], inp["lf0"],
                inp["lf0_lengths"])
    else:
        args = (inp["x"], inp["x_lengths"], y, y_lengths)
    torch.manual_seed(3); np.random.seed(3)
    ours = model.compute_loss(*args, out_size=None)
    torch.manual_seed(3); np.random.seed(3)
    theirs = ref_model.compute_loss(*args, out_size=None)
    assert len(ours) == len(theirs) == (4 if variant == "dex" else 3)
    for a, b in zip(ours, theirs):
        assert torch.isfinite(a).all() and torch.allclose(a, b, rtol=1e-5, atol=1e-6), (float(a), float(b))
    sum(ours[:3]).backward()
    grads = [p.grad for p in model.decoder.parameters() if p.grad is not None]
    assert len(grads) > 50 and all(torch.isfinite(g).all() for g in grads)
    assert list(model.state_dict().keys()) == list(ref_model.state_dict().keys())     # the twin adds no keys
    RT.set_reference_dir(None)


def test_training_entry_point_without_a_reference_checkout_says_what_to_do(tmp_path, monkeypatch):
    import dexb200.model.reference_twin as RT
    RT.set_reference_dir(None)
    monkeypatch.delenv("DEXB_REFERENCE_DIR", raising=False)
    monkeypatch.chdir(tmp_path)
    model, _ = build("gedex")
    with pytest.raises(RuntimeError, match="DEXB_REFERENCE_DIR"):
        model.compute_loss(torch.zeros(1, 4, dtype=torch.long), torch.tensor([4]), torch.zeros(1, 80, 8), torch.tensor([8]))


def test_dropin_package_serves_the_reference_import_lines():
    """`from model import DeXTTS` / `from model.utils import fix_len_compatibility` (DEX-TTS/synthesize.py:11, main.py:14) resolve to the
    CUDA package when dex-tts_b200/dropin is first on the path -- in a fresh interpreter, so no other `model` is cached."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(root, "dex-tts_b200", "dropin"), os.path.join(root, "dex-tts_b200")]))
    code = ("from model import DeXTTS, GeDEXTTS\n"
            "from model.utils import fix_len_compatibility, sequence_mask\n"
            "from model.diffusion import Diffusion\n"
            "import dexb200.model as M\n"
            "assert DeXTTS is M.DeXTTS and GeDEXTTS is M.GeDEXTTS and Diffusion is M.Diffusion\n"
            "assert fix_len_compatibility(61) == 64\n"
            "print('ok')\n")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
