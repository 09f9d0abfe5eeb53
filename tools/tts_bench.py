"""Time the whole drop-in model (dexb200.model.DeXTTS.forward: style stage -> text encoder -> duration / alignment glue -> N-step reverse
diffusion) and its stages at the shapes of BASELINE.json's C2:  python tools/tts_bench.py [B] [Tx] [Ts] [n_timesteps] [iters]
-> one JSON line.  CUDA events on the current stream; the forward contains the path's one host sync (the predicted mel length)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "dex-tts_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
from dexb200.model import align_durations, style_fusion                  # noqa: E402
from dexb200.synth import synth_lf0, synth_ref_mel, synth_text            # noqa: E402
from test_tts_module_cpu import build                                     # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
Tx = int(sys.argv[2]) if len(sys.argv) > 2 else 128
Ts = int(sys.argv[3]) if len(sys.argv) > 3 else 259
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 50
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5

model, _ = build("dex")
model = model.cuda().eval()
txt = synth_text(B, Tx, seed=1)
mel = synth_ref_mel(B, Ts, seed=2)
lf0 = synth_lf0(B, Ts, seed=3)
x, xl = txt["x"].cuda(), txt["x_lengths"].cuda()
ref, rl = mel["ref"].cuda(), mel["ref_lengths"].cuda()
f0 = lf0["lf0"].cuda()
mask = mel["mask"].cuda()


def timed(fn, n=iters):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, out


def style():
    le, ld = model.lf0_encoder(f0, mask)
    zb, zd, _ = model.tv_encoder(ref, mask)
    se, sd = style_fusion(model.conv_sty, zb, zd, mask, le, ld, mask)
    _, skips = model.tiv_encoder(ref, mask)
    return se, sd, skips


ms_style, (sty_enc, sty_dec, skips) = timed(style, 20)
ms_text, (mu_x, logw, x_mask) = timed(lambda: model.encoder(x, xl, sty_enc), 20)
ms_align, (mu_y, y_mask, attn, y_len, y_max) = timed(lambda: align_durations(logw, x_mask, mu_x), 20)
ms_loop, _ = timed(lambda: model.decoder(mu_y, y_mask, mu_y, skips, rl, sty_dec, rl, n_timesteps=steps, infer=True, temperature=1.5))
ms_all, (enc_out, dec_out, _) = timed(lambda: model(x, xl, ref, rl, ref, rl, f0, rl, n_timesteps=steps, temperature=1.5))
frames = int(y_len.sum())
print(json.dumps({"what": "DeXTTS.forward (whole drop-in model)", "B": B, "Tx": Tx, "Ts": Ts, "n_timesteps": steps, "Ty_padded": int(mu_y.shape[-1]),
                  "mel_frames": frames, "ms_style_stage": round(ms_style, 3), "ms_text_encoder": round(ms_text, 3),
                  "text_encoder_launches": model.encoder.cuda_engine().launches, "ms_align": round(ms_align, 3),
                  "ms_loop": round(ms_loop, 2), "ms_forward": round(ms_all, 2), "mel_frames_per_s": round(frames / ms_all * 1e3, 1),
                  "finite": bool(torch.isfinite(dec_out).all())}))
