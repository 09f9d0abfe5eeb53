"""``audio.stft.TacotronSTFT`` on the CUDA STFT kernel.  Replaces DEX-TTS/audio/stft.py:130-178.

The reference builds a dense windowed DFT basis and runs it as a ``conv1d`` on the GPU, then moves the result back to the CPU
(stft.py:64-69); here ``mel_spectrogram`` is one launch of ``dexb_stft_mel`` (FFT + triangular mel product + log, csrc/kernels_stft.cu)
and returns CPU tensors like upstream.  Third-party arithmetic restated (neither librosa nor scipy is needed at run time):
  * ``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`` of the pinned librosa 0.9.2 (requirements.txt:19; htk=False,
    norm='slaney'): Slaney mel scale (linear below 1 kHz, log above), triangles between neighbouring band edges, area-normalised;
  * ``scipy.signal.get_window('hann', win_length, fftbins=True)`` = the periodic Hann window, centre-padded to ``filter_length``.
"""
import numpy as np
import torch

from ..engine import stft_mel
from .audio_processing import dynamic_range_compression, dynamic_range_decompression


def _slaney_hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f * 3.0 / 200.0
    log_region = 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) * (27.0 / np.log(6.4))
    return np.where(f >= 1000.0, log_region, lin)


def _slaney_mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= 15.0, 1000.0 * np.exp((m - 15.0) * (np.log(6.4) / 27.0)), m * 200.0 / 3.0)


def slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax):
    """(n_mels, 1 + n_fft // 2) float32, the matrix librosa 0.9.2 ``filters.mel`` returns for these arguments."""
    bins = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    edges = _slaney_mel_to_hz(np.linspace(_slaney_hz_to_mel(fmin), _slaney_hz_to_mel(fmax), n_mels + 2))
    width = np.diff(edges)
    rising = (bins[None, :] - edges[:-2, None]) / width[:-1, None]
    falling = (edges[2:, None] - bins[None, :]) / width[1:, None]
    tri = np.clip(np.minimum(rising, falling), 0.0, None)
    tri *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return tri.astype(np.float32)


class TacotronSTFT(torch.nn.Module):
    def __init__(self, filter_length, hop_length, win_length, n_mel_channels, sampling_rate, mel_fmin, mel_fmax):
        super().__init__()
        if filter_length != 1024 or hop_length != 256:
            raise NotImplementedError("dexb_stft_mel is instantiated for filter_length 1024 / hop_length 256 (every shipped config)")
        if win_length > filter_length:
            raise ValueError("win_length must not exceed filter_length")
        self.n_mel_channels, self.sampling_rate = n_mel_channels, sampling_rate
        self.filter_length, self.hop_length, self.win_length = filter_length, hop_length, win_length
        n = np.arange(win_length, dtype=np.float64)
        win = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)
        lpad = (filter_length - win_length) // 2
        win = np.pad(win, (lpad, filter_length - win_length - lpad))
        self.register_buffer("window", torch.from_numpy(win).float())
        self.register_buffer("mel_basis", torch.from_numpy(slaney_mel_basis(sampling_rate, filter_length, n_mel_channels, mel_fmin, mel_fmax)))

    def spectral_normalize(self, magnitudes):
        return dynamic_range_compression(magnitudes)

    def spectral_de_normalize(self, magnitudes):
        return dynamic_range_decompression(magnitudes)

    @torch.no_grad()
    def mel_spectrogram(self, y):
        """y (B, T) in [-1, 1] -> (log-mel (B, n_mel_channels, 1 + T // hop), energy (B, 1 + T // hop)), both on the CPU like upstream."""
        assert torch.min(y.data) >= -1
        assert torch.max(y.data) <= 1
        if not torch.cuda.is_available():
            raise RuntimeError("audio.stft.TacotronSTFT runs on CUDA (sm_100a) only; there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
        mel, energy = stft_mel(y.to(dev), self.window.to(dev), self.mel_basis.to(dev), n_fft=self.filter_length, hop=self.hop_length,
                               return_energy=True)
        return mel.cpu(), energy.cpu()
