"""``audio.audio_processing`` -- the two element-wise helpers of DEX-TTS/audio/audio_processing.py:85-100 that the mel path names.
(Griffin-Lim / window_sumsquare belong to the inverse STFT, which no entry script of the reference calls.)"""
import torch


def dynamic_range_compression(x, C=1, clip_val=1e-5):
    return torch.log(torch.clamp(x, min=clip_val) * C)


def dynamic_range_decompression(x, C=1):
    return torch.exp(x) / C
