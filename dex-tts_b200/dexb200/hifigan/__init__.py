"""``hifigan`` -- the vocoder package of the reference (``import hifigan``, DEX-TTS/src/utils.py:10,251-281) on the CUDA path."""
from .models import Generator  # noqa: F401


class AttrDict(dict):
    """DEX-TTS/hifigan/__init__.py:4-7."""

    def __init__(self, *args, **kwargs):
        super(AttrDict, self).__init__(*args, **kwargs)
        self.__dict__ = self
