"""Time the CUDA TIV encoder (dexb_tiv_forward) on synthetic reference mels: python tools/tiv_bench.py [B] [T] [iters]."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dex-tts_b200"))
from dexb200.model import TIVEncoder                      # noqa: E402
from dexb200.synth import synth_ref_mel, synth_tiv_weights   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 259
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
m = TIVEncoder(c_in=80, c_out=64, num_layer=6, c_h=128)
m.load_state_dict(synth_tiv_weights(prefix=""), strict=True)
m = m.cuda().eval()
inp = synth_ref_mel(B, T, seed=1, ragged=False)
ref, mask = inp["ref"].cuda(), inp["mask"].cuda()
for _ in range(5):
    m(ref, mask)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(iters):
    m(ref, mask)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / iters
flop = 2.0 * B * T * 3 * (80 * 128 + 12 * 128 * 128 + 128 * 64)      # 14 conv1d, k = 3 (algorithmic, 2*MAC)
print(json.dumps({"what": "TIV encoder forward (14 conv1d + BN/ReLU/InstanceNorm1D fusions)", "B": B, "T": T, "ms": ms,
                  "launches": m.cuda_engine().launches, "gflop": flop / 1e9, "tflops": flop / ms / 1e9,
                  "utterances_per_s": B / ms * 1e3}))
