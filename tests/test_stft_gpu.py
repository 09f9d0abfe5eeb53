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


def test_energy_output_and_ragged_frame_counts():
    """energy = ||magnitude||_2 per frame (stft.py:176) and frame counts that are not a multiple of the 8 frames a CTA produces;
    odd utterance strides exercise the unaligned staging path."""
    from dexb200.engine import stft_mel
    g = np.random.default_rng(5)
    for B, S in ((3, 6001), (2, 66150), (1, 513), (5, 2303)):
        wav = g.uniform(-1.0, 1.0, size=(B, S)).astype(np.float32)
        win = torch.from_numpy(SO.hann_periodic(1024).astype(np.float32)).cuda()
        fb = torch.from_numpy(SO.mel_filterbank()).cuda()
        mel, en = stft_mel(torch.from_numpy(wav).cuda(), win, fb, return_energy=True)
        ref = SO.mel_spectrogram(wav)
        assert mel.shape == ref.shape == (B, 80, S // 256 + 1)
        assert np.abs(mel.cpu().numpy() - ref).max() < 3e-4
        x = np.pad(wav.astype(np.float64), ((0, 0), (512, 512)), mode="reflect")
        idx = np.arange(1024)[None, :] + 256 * np.arange(S // 256 + 1)[:, None]
        mag = np.abs(np.fft.rfft(x[:, idx] * SO.hann_periodic(1024)[None, None, :], axis=-1))
        en_ref = np.sqrt((mag ** 2).sum(-1))
        assert np.abs(en.cpu().numpy() - en_ref).max() / en_ref.max() < 1e-5


def test_audio_dropin_matches_reference_fixture():
    """``Audio.stft.TacotronSTFT`` / ``Audio.tools.get_mel_from_wav`` of the drop-in package (the calls of synthesize.py:49,79-85)
    against the outputs of the reference's own classes on the first 1.2 s of syn_samples/sample1.wav."""
    from dexb200.audio.stft import TacotronSTFT
    from dexb200.audio.tools import get_mel_from_wav
    stft = TacotronSTFT(1024, 256, 1024, 80, 22050, 0.0, 8000.0)      # config/VCTK/base.yaml preprocess block
    g = np.load([p for p in GOLD if "sample1" in p][0])
    w = g["wav"]
    wav = (w.astype(np.float32) / 32768.0) if w.dtype == np.int16 else w
    for i in range(wav.shape[0]):
        mel, energy = get_mel_from_wav(wav[i], stft)
        assert mel.dtype == np.float32 and mel.shape == g["mel"][i].shape and energy.shape == (mel.shape[1],)
        live = g["mel"][i] > np.log(1e-5) + 1e-3
        assert np.abs(mel - g["mel"][i])[live].max() < 3e-4
