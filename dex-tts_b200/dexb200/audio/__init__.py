"""``audio`` -- the reference-audio front-end of the reference (``import audio as Audio``, DEX-TTS/synthesize.py:15) with the
STFT -> mel -> log chain on the CUDA kernel ``dexb_stft_mel``:

    Audio.stft.TacotronSTFT(...).mel_spectrogram(y)     DEX-TTS/audio/stft.py:130-178   (synthesize.py:79-85)
    Audio.tools.get_mel_from_wav(wav, STFT)             DEX-TTS/audio/tools.py:8-15     (synthesize.py:49)

Same constructor arguments, same return values (CPU tensors / numpy arrays, as upstream's ``.cpu()`` / ``.numpy()`` produce).
No CPU fallback: without the CUDA extension / a CUDA device the call raises."""
from . import audio_processing, stft, tools  # noqa: F401
