"""Generate tests/golden/stft_*.npz by running the reference's own STFT / TacotronSTFT classes
(DEX-TTS/audio/stft.py, imported unmodified from /root/reference) in the build container.

Shims (no reference file is edited): ``librosa`` is not installed, so a stub module supplies ``util.pad_center`` /
``util.tiny`` and ``filters.mel`` (the latter = oracle/stft_oracle.mel_filterbank, cross-checked against torchaudio in
tests/test_stft_oracle.py); ``Tensor.cuda`` is patched to the identity because STFT.transform hard-codes ``.cuda()``
(stft.py:68-69) and this container has no GPU.  Inputs: seeded synthetic audio, and the first 1.2 s of the reference's
only audio fixture, DEX-TTS/syn_samples/sample1.wav.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import stft_oracle as SO  # noqa: E402

REF = os.environ.get("DEX_REFERENCE_ROOT", "/root/reference")


def load_reference_audio():
    lib = types.ModuleType("librosa")
    util = types.ModuleType("librosa.util")
    filters = types.ModuleType("librosa.filters")

    def pad_center(data, size, axis=-1, **kw):
        n = data.shape[axis]
        lpad = int((size - n) // 2)
        lengths = [(0, 0)] * data.ndim
        lengths[axis] = (lpad, int(size - n - lpad))
        return np.pad(data, lengths, **kw)

    util.pad_center = pad_center
    util.tiny = lambda x: np.finfo(np.float32).tiny
    util.normalize = lambda x, **kw: x
    filters.mel = lambda sr, n_fft, n_mels, fmin, fmax: SO.mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    lib.util, lib.filters = util, filters
    sys.modules.update({"librosa": lib, "librosa.util": util, "librosa.filters": filters})
    sys.path.insert(0, os.path.join(REF, "DEX-TTS"))
    import audio.stft as ref_stft
    return ref_stft


def main():
    ref_stft = load_reference_audio()
    torch.Tensor.cuda = lambda self, *a, **k: self
    tac = ref_stft.TacotronSTFT(1024, 256, 1024, 80, 22050, 0, 8000)     # DEX-TTS/synthesize.py:79-85
    g = np.random.default_rng(2024)
    cases = {}
    # band-limited-ish noise, two lengths (one not a multiple of the hop), amplitude inside [-1, 1]
    for name, B, S in (("stft_noise", 2, 6000), ("stft_short", 1, 1500)):
        x = g.uniform(-0.5, 0.5, size=(B, S)).astype(np.float32)
        x[:, ::7] *= 0.1
        cases[name] = x
    from scipy.io import wavfile
    sr, wav = wavfile.read(os.path.join(REF, "DEX-TTS", "syn_samples", "sample1.wav"))
    assert sr == 22050
    w = (wav[: int(1.2 * sr)].astype(np.float32) / 32768.0)[None]
    cases["stft_sample1"] = w
    for name, x in cases.items():
        with torch.no_grad():
            mel, energy = tac.mel_spectrogram(torch.from_numpy(x))
        out = dict(wav=x if name != "stft_sample1" else (x * 32768.0).astype(np.int16), mel=mel.numpy().astype(np.float32))
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **out)
        print(name, x.shape, "->", tuple(mel.shape), f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
