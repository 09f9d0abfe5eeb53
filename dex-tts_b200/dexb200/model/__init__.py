"""``from model import DeXTTS`` / ``GeDEXTTS`` (DEX-TTS/model/__init__.py:1, GeDEX-TTS/model/__init__.py:1) and drop-in replacements
for the reference's ``model.diffusion.Diffusion`` (DEX-TTS and GeDEX-TTS flavours) and, in front of the
loop, ``model.ref_encoder.TIVEncoder`` / ``TVEncoder`` / ``LF0Encoder``, the style fusion, ``model.text_encoder.TextEncoder`` and the duration /
alignment glue of ``DeXTTS.forward`` (``model.utils``)."""
from .diffusion import Diffusion, GeDiffusion  # noqa: F401
from .ref_encoder import LF0Encoder, TIVEncoder, TVEncoder, style_fusion  # noqa: F401
from .utils import align_durations, fix_len_compatibility, sequence_mask  # noqa: F401
from .text_encoder import GeTextEncoder, TextEncoder  # noqa: F401
from .tts import DeXTTS, GeDEXTTS  # noqa: F401
from . import monotonic_align  # noqa: F401,E402
