"""Drop-in replacements for the reference's ``model.diffusion.Diffusion`` (DEX-TTS and GeDEX-TTS flavours)."""
from .diffusion import Diffusion, GeDiffusion  # noqa: F401
