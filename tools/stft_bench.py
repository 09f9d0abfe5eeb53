"""Device time of dexb_stft_mel at BASELINE config 3 (B = 32 utterances of 3 s = 66 150 samples) against the HBM roofline.
Algorithmic bytes: 4 S read + 4 (n_mels [+ 1 energy]) (1 + S / 256) written per utterance."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "dex-tts_b200")]
from dexb200.audio.stft import slaney_mel_basis  # noqa: E402
from dexb200.engine import stft_mel  # noqa: E402


def main():
    B, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32, 66150)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    g = torch.Generator().manual_seed(1)
    # 16 distinct input batches (> L2 is not reachable at 8.5 MB per batch: rotate buffers and flush L2 between timed launches instead)
    wavs = [(torch.rand(B, S, generator=g) - 0.5).cuda() for _ in range(4)]
    win = torch.hann_window(1024, periodic=True).cuda()
    fb = torch.from_numpy(slaney_mel_basis(22050, 1024, 80, 0.0, 8000.0)).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for w in wavs:
        stft_mel(w, win, fb)
    torch.cuda.synchronize()
    times = []
    for it in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        stft_mel(wavs[it % 4], win, fb)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    times.sort()
    ms = times[len(times) // 2]
    frames = S // 256 + 1
    nbytes = B * (4 * S + 4 * 80 * frames)
    print(json.dumps({"kernel": "k_stft_mel", "B": B, "S": S, "frames": frames, "ms_median": ms, "ms_min": times[0],
                      "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / (ms * 1e-3) / 1e9, "peak_gbs": peak,
                      "frac": nbytes / (ms * 1e-3) / 1e9 / peak, "utterances_per_s": B / (ms * 1e-3),
                      "note": "cold L2 (256 MB flush between launches); includes the per-call torch.empty of the outputs"}))


if __name__ == "__main__":
    main()
