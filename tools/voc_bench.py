"""Device-timed vocoder forward (dexb_voc_forward): python tools/voc_bench.py [B] [T]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dex-tts_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import json  # noqa: E402

import torch  # noqa: E402

import vocoder_oracle as V  # noqa: E402
from dexb200.hifigan.models import VocoderEngine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 512
cfg = dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
           resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3)
eng = VocoderEngine(cfg)
eng.load_state_dict(V.synth_vocoder_weights())
mel = (torch.randn(B, 80, T) * 1.5 - 4.0).cuda()
for _ in range(3):
    wav = eng.forward(mel)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
a.record()
for _ in range(n):
    wav = eng.forward(mel)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / n
# algorithmic MACs per output configuration (2 live taps of each ConvTranspose)
macs, ch, L = 80 * 512 * 7 * T, 512, T
for u in (8, 8, 2, 2):
    macs += L * u * ch * (ch // 2) * 2
    L *= u; ch //= 2
    macs += sum(L * ch * ch * k * 6 for k in (3, 7, 11))
macs += L * ch * 7
sec = B * T * 256 / 22050
print(json.dumps({"B": B, "T": T, "ms": ms, "launches": eng.launches, "audio_s": sec, "rtf": ms / 1e3 / sec,
                  "algorithmic_tflops": 2 * macs * B / ms / 1e9, "finite": bool(torch.isfinite(wav).all())}))
