from dexb200.audio.tools import get_mel_from_wav  # noqa: F401
