"""A package literally named ``audio`` (``import audio as Audio``, DEX-TTS/synthesize.py:15): with ``dex-tts_b200/dropin`` in front of
the reference checkout on ``sys.path`` the unmodified ``synthesize.py`` extracts its reference mel with the CUDA STFT kernel.
Everything is re-exported from ``dexb200.audio``."""
from . import audio_processing, stft, tools  # noqa: F401
