"""Parity of the CUDA STFT -> mel -> log kernel (dexb_stft_mel through the C ABI) against the reference's TacotronSTFT
outputs (golden fixtures) and the CPU oracle at the BASELINE config-3 size (3 s of audio, batch 32)."""
import glob
import os

import numpy as np
import pytest
import torch

import stft_oracle as SO

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "stft_*.npz")))


def run_gpu(wav):
    from dexb200.engine import stft_mel
    win = torch.from_numpy(SO.hann_periodic(1024).astype(np.float32)).cuda()
    fb = torch.from_numpy(SO.mel_filterbank()).cuda()
    return stft_mel(torch.from_numpy(np.ascontiguousarray(wav)).cuda(), win, fb).cpu().numpy()


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_stft_kernel_matches_reference_golden(path):
    g = np.load(path)
    w = g["wav"]
    wav = (w.astype(np.float32) / 32768.0) if w.dtype == np.int16 else w
    mel = run_gpu(wav)
    ref = g["mel"]
    assert mel.shape == ref.shape
    live = ref > np.log(1e-5) + 1e-3
    assert np.abs(mel - ref)[live].max() < 3e-4          # log-mel, absolute (= 3e-4 relative on the mel energy)
    assert np.abs(mel - ref).max() < 5e-3


def test_stft_kernel_config3_size():
    """B = 32 utterances of 3 s (66 150 samples -> 259 frames): CUDA vs float64 oracle."""
    g = np.random.default_rng(3)
    wav = g.uniform(-0.5, 0.5, size=(32, 66150)).astype(np.float32)
    mel = run_gpu(wav)
    ref = SO.mel_spectrogram(wav)
    assert mel.shape == (32, 80, 259)
    assert np.abs(mel - ref).max() < 3e-4
    # silence maps to the clamp floor exactly
    z = run_gpu(np.zeros((1, 4096), dtype=np.float32))
    assert np.allclose(z, np.log(1e-5))
