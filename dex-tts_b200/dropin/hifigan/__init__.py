"""A package literally named ``hifigan`` (``import hifigan``, DEX-TTS/src/utils.py): with ``dex-tts_b200/dropin`` in front of the
reference checkout on ``sys.path`` the unmodified ``get_vocoder`` builds the CUDA generator.  Re-exported from ``dexb200.hifigan``."""
from dexb200.hifigan import AttrDict, Generator  # noqa: F401
from . import models  # noqa: F401
