"""dexb200 -- B200-native reverse-diffusion path of DEX-TTS / GeDEX-TTS."""
from .manifest import DecoderCfg, decoder_manifest  # noqa: F401
