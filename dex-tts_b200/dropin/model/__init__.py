"""A package literally named ``model``: put ``dex-tts_b200/dropin`` (and ``dex-tts_b200``) in front of the reference checkout on
``sys.path`` / ``PYTHONPATH`` and the reference's entry scripts run on the CUDA path without an edit --

    from model import DeXTTS                          DEX-TTS/synthesize.py:11, src/evaluation.py:15
    from model import GeDEXTTS                        GeDEX-TTS/synthesize.py:9, src/evaluation.py:15
    from model.utils import fix_len_compatibility     DEX-TTS/main.py:14, src/dataset.py:10 (GeDEX-TTS/main.py:9)

Everything is re-exported from ``dexb200.model``; nothing is implemented here.  (``model.augmentation``, imported by the training
data loader src/dataset.py:11, is training-only and not provided.)"""
from dexb200.model import *  # noqa: F401,F403
from dexb200.model import DeXTTS, GeDEXTTS  # noqa: F401
