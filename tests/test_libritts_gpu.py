"""DEX-TTS at the LibriTTS sizes (DEX-TTS/config/LibriTTS/base.yaml) end to end on the GPU through the drop-in ``DeXTTS``: against the
fixtures of the unmodified reference model (tests/golden/libritts_dex_*.npz) and, for the decoder, the CPU oracle on the GPU's own
conditioning.  Kernels that only run at these sizes: conv_in / final kernel at 128 channels, 256-channel GroupNorm with per-tile
statistics, LinearAttention context on the CUDA cores (C = 256), TV adaptor and DiT attention through GEMM -> softmax -> GEMM
(256 channels / head dim 192), positional convolution with four taps per K chunk (48 channels per group), GRU with 128 hidden units."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

import dex_oracle as O
from dexb200.synth import seeded_noise
from parity import REL_TOL, per_bin_violation, tensor_rel_err
from test_libritts import DCFG, build_libritts

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from make_golden_tts import synth_tts_inputs  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "libritts_dex_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_libritts_model_forward_on_gpu(path, monkeypatch):
    g = np.load(path)
    B, Tx, Ts, steps, ragged, seed = [int(v) for v in g["meta"]]
    temperature, length_scale = [float(v) for v in g["scale"]]
    inp = {k: v.cuda() for k, v in synth_tts_inputs("dex", B, Tx, Ts, seed, bool(ragged)).items() if torch.is_tensor(v)}
    model, w = build_libritts()
    model = model.cuda().eval()
    cap = {}
    model.decoder.register_forward_pre_hook(lambda m, args, kwargs: cap.update(args=args, kwargs=kwargs), with_kwargs=True)
    noise = seeded_noise(seed + 3)
    drawn = []
    real_randn = torch.randn

    def fake_randn(*shape, **kw):
        shp = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
        z = noise(shp)
        drawn.append(z)
        return z.to(kw.get("device", "cpu"))
    monkeypatch.setattr(torch, "randn", fake_randn)
    enc_out, dec_out, attn = model(inp["x"], inp["x_lengths"], inp["ref"], inp["ref_lengths"], inp["ref"], inp["ref_lengths"],
                                   inp["lf0"], inp["lf0_lengths"], spk=None, n_timesteps=steps, temperature=temperature,
                                   length_scale=length_scale)
    torch.cuda.synchronize()
    monkeypatch.setattr(torch, "randn", real_randn)
    assert len(drawn) == 1 and torch.isfinite(dec_out).all()
    a = cap["args"]
    mu_y, y_mask = a[0].cpu(), a[1].cpu()
    cond = dict(sty=a[5].cpu(), sty_lengths=a[6].cpu(), ref_skips=[r.cpu() for r in a[3]])
    with torch.no_grad():
        y = O.reverse_diffusion(w, O.make_cfg("dex", **DCFG), drawn[0], y_mask, mu_y, steps, temperature, cond)
    v_dec = per_bin_violation(dec_out.cpu(), y[:, :, :dec_out.shape[-1]])
    print(f"{os.path.basename(path)}: decoder vs oracle on the same conditioning {v_dec:.2e}, simt fallback GEMMs "
          f"{model.decoder.cuda_engine().simt_fallbacks}")
    assert v_dec < REL_TOL
    shape = tuple(int(n) for n in g["attn_shape"])
    attn_ref = np.unpackbits(g["attn"], axis=-1, count=shape[-1]).astype(np.float32).reshape(shape)
    assert tuple(attn.shape) == shape and np.array_equal(attn.cpu().numpy(), attn_ref)
    v_enc = per_bin_violation(enc_out.cpu(), torch.from_numpy(g["enc_out"]))
    e_dec = tensor_rel_err(dec_out.cpu(), torch.from_numpy(g["dec_out"]))
    print(f"{os.path.basename(path)}: alignment identical, enc_out {v_enc:.2e}, dec_out vs the reference fixture {e_dec:.2e}")
    assert v_enc < REL_TOL and e_dec < REL_TOL
