"""Drop-in replacements for the reference's ``model.diffusion.Diffusion`` (DEX-TTS and GeDEX-TTS flavours) and, in front of the
loop, ``model.ref_encoder.TIVEncoder`` / ``TVEncoder`` / ``LF0Encoder`` and the style fusion of ``DeXTTS.forward``."""
from .diffusion import Diffusion, GeDiffusion  # noqa: F401
from .ref_encoder import LF0Encoder, TIVEncoder, TVEncoder, style_fusion  # noqa: F401
