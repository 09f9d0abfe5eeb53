"""Pin the chain of stage oracles (oracle/tts_oracle.py: style stage -> text encoder -> duration / alignment glue -> reverse diffusion,
in the order of DeXTTS.forward, DEX-TTS/model/tts.py:33-74, and GeDEXTTS.forward, GeDEX-TTS/model/tts.py:27-56) against the outputs
(enc_out, dec_out, attn) of the unmodified reference models' own ``forward`` (tests/golden/tts_*.npz, made by oracle/make_golden_tts.py
in the build container from the reference's config yaml) -- the signature SURVEY.md section 8b asks a drop-in to keep."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

import tts_oracle as TT
from dexb200.synth import seeded_noise, synth_tts_weights
from parity import REL_TOL, per_bin_violation

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from make_golden_tts import synth_tts_inputs  # noqa: E402  (imports ref_loader, which only touches /root/reference when called)

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "tts_*.npz")))


def test_golden_present():
    assert len(GOLD) >= 3


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_forward_oracle_matches_reference_forward(path):
    g = np.load(path)
    variant = str(g["variant"])
    B, Tx, Ts, steps, ragged, seed = [int(v) for v in g["meta"]]
    temperature, length_scale = [float(v) for v in g["scale"]]
    inp = synth_tts_inputs(variant, B, Tx, Ts, seed, bool(ragged))
    w = synth_tts_weights(variant)
    with torch.no_grad():
        if variant == "dex":
            enc_out, dec_out, attn = TT.dextts_forward(w, inp["x"], inp["x_lengths"], inp["ref"], inp["ref_lengths"], inp["ref"],
                                                       inp["ref_lengths"], inp["lf0"], inp["lf0_lengths"], steps, seeded_noise(seed + 3),
                                                       temperature, length_scale)
        else:
            enc_out, dec_out, attn = TT.gedextts_forward(w, inp["x"], inp["x_lengths"], steps, seeded_noise(seed + 3), temperature,
                                                         length_scale)
    shape = tuple(int(v) for v in g["attn_shape"])
    attn_ref = np.unpackbits(g["attn"], axis=-1, count=shape[-1]).astype(np.float32).reshape(shape)
    assert attn.shape == shape and np.array_equal(attn.numpy(), attn_ref)          # the hard alignment: bit-exact
    assert enc_out.shape == g["enc_out"].shape and dec_out.shape == g["dec_out"].shape
    v_enc = per_bin_violation(enc_out, torch.from_numpy(g["enc_out"]))
    v_dec = per_bin_violation(dec_out, torch.from_numpy(g["dec_out"]))
    print(f"{os.path.basename(path)}: enc_out {v_enc:.2e} dec_out {v_dec:.2e}")
    # dec_out through the whole chain: the restated GRU of the LF0 encoder moves `sty` by ~1e-6 of its RMS (reassociation), and two
    # sampler steps of the randomly initialised decoder amplify that up to ~100x (measured 2.3e-4 on tts_dex_b2r, 1.3e-5 on b1) --
    # so the chain is held to the path tolerance here and the decoder is pinned tightly on the reference's own `sty` below.
    assert v_enc < 2e-5 and v_dec < REL_TOL


@pytest.mark.parametrize("path", [p for p in GOLD if "tts_dex" in p], ids=[os.path.basename(p)[:-4] for p in GOLD if "tts_dex" in p])
def test_decoder_oracle_on_the_reference_style_input(path):
    """Same fixtures, but the loop's `sty` is the tensor the reference itself fed to its decoder (stored in the fixture), mu_y is the
    reference's enc_out zero-padded to fix_len_compatibility, ref_skips come from the TIV oracle (bit-identical to the reference):
    the remaining difference is the decoder oracle alone (measured 0.0 in the build container)."""
    import dex_oracle as O
    g = np.load(path)
    B, Tx, Ts, steps, ragged, seed = [int(v) for v in g["meta"]]
    temperature, length_scale = [float(v) for v in g["scale"]]
    inp = synth_tts_inputs("dex", B, Tx, Ts, seed, bool(ragged))
    w = synth_tts_weights("dex")
    enc_out = torch.from_numpy(g["enc_out"])
    y_max = enc_out.shape[-1]
    Ty = -(-y_max // 4) * 4
    mu_y = torch.nn.functional.pad(enc_out, (0, Ty - y_max))
    y_lengths = torch.from_numpy(np.unpackbits(g["attn"], axis=-1, count=int(g["attn_shape"][-1]))
                                 .reshape(tuple(int(v) for v in g["attn_shape"])).astype(np.int64)).sum((1, 2, 3)).clamp_min(1)
    y_mask = TT._seq_mask(y_lengths, Ty)
    with torch.no_grad():
        _, ref_skips = O.tiv_encoder(w, inp["ref"], TT._seq_mask(inp["ref_lengths"], Ts))
        cond = dict(sty=torch.from_numpy(g["sty_dec"]), sty_lengths=inp["ref_lengths"], ref_skips=ref_skips)
        y = O.reverse_diffusion(w, O.make_cfg("dex"), seeded_noise(seed + 3)((B, 80, Ty)), y_mask, mu_y, steps, temperature, cond)
    v = per_bin_violation(y[:, :, :y_max], torch.from_numpy(g["dec_out"]))
    print(f"{os.path.basename(path)}: decoder oracle on the reference's sty: {v:.2e}")
    assert v < 2e-5
