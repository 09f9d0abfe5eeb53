"""Drop-in replacements for the reference's ``model.diffusion.Diffusion`` (DEX-TTS and GeDEX-TTS flavours) and, in front of the
loop, ``model.ref_encoder.TIVEncoder`` / ``TVEncoder``."""
from .diffusion import Diffusion, GeDiffusion  # noqa: F401
from .ref_encoder import TIVEncoder, TVEncoder  # noqa: F401
