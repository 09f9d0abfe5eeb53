"""``audio.tools.get_mel_from_wav`` -- DEX-TTS/audio/tools.py:8-15 (called at synthesize.py:49 and by the pre-processing scripts)."""
import numpy as np
import torch


def get_mel_from_wav(audio, _stft):
    """audio: 1-D float array -> (mel (n_mels, frames), energy (frames,)) float32 numpy arrays; samples are clipped to [-1, 1] first."""
    wav = torch.clip(torch.FloatTensor(audio).unsqueeze(0), -1, 1)
    melspec, energy = _stft.mel_spectrogram(wav)
    return melspec.squeeze(0).numpy().astype(np.float32), energy.squeeze(0).numpy().astype(np.float32)
