"""CPU oracle for the reference-audio feature front-end (STFT -> mel -> log).

TEST INFRASTRUCTURE ONLY (see oracle/dex_oracle.py header for who may import this).

Restates  TacotronSTFT.mel_spectrogram  DEX-TTS/audio/stft.py:159-178
          STFT.transform                DEX-TTS/audio/stft.py:52-81  (reflect pad n_fft/2; Hann-windowed DFT basis; hop stride)
          dynamic_range_compression     DEX-TTS/audio/audio_processing.py:85-91
Third-party arithmetic on the path (absent from /root/reference, restated from the published algorithms):
  * librosa==0.9.2 (DEX-TTS/requirements.txt:19) ``filters.mel(sr, n_fft, n_mels, fmin, fmax)`` -- Slaney mel scale,
    Slaney area normalisation (call site stft.py:145-147) and ``util.pad_center`` (stft.py:42; a no-op for win == n_fft);
  * scipy ``get_window('hann', N, fftbins=True)`` = periodic Hann (stft.py:41).
Parity pin: tests/golden/stft_*.npz hold outputs of the reference's own STFT/TacotronSTFT classes run in the build
container (oracle/make_golden_stft.py; librosa is stubbed there with THIS file's mel filter, which is itself
cross-checked against torchaudio.functional.melscale_fbanks(norm='slaney', mel_scale='slaney') in
tests/test_stft_oracle.py).
"""
import numpy as np


def hann_periodic(n):
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)).astype(np.float64)


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(sr=22050, n_fft=1024, n_mels=80, fmin=0.0, fmax=8000.0):
    """librosa.filters.mel (0.9.2 defaults: htk=False, norm='slaney', dtype float32) -> (n_mels, 1 + n_fft // 2)."""
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0, float(sr) / 2, n_bins, endpoint=True)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights.astype(np.float32)


def mel_spectrogram(wav, n_fft=1024, hop=256, mel_basis=None, window=None, clip_val=1e-5):
    """wav (B, S) float in [-1, 1] -> log-mel (B, n_mels, 1 + S // hop) float32 (float64 arithmetic inside)."""
    wav = np.asarray(wav, dtype=np.float64)
    assert wav.min() >= -1 and wav.max() <= 1                      # stft.py:169-170
    if mel_basis is None:
        mel_basis = mel_filterbank(n_fft=n_fft)
    if window is None:
        window = hann_periodic(n_fft)
    B, S = wav.shape
    pad = n_fft // 2
    x = np.pad(wav, ((0, 0), (pad, pad)), mode="reflect")
    n_frames = (x.shape[1] - n_fft) // hop + 1
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = x[:, idx] * window[None, None, :]                     # (B, frames, n_fft)
    spec = np.fft.rfft(frames, axis=-1)
    mag = np.abs(spec)                                             # (B, frames, bins)
    mel = np.einsum("mk,bfk->bmf", mel_basis.astype(np.float64), mag)
    return np.log(np.maximum(mel, clip_val)).astype(np.float32)
