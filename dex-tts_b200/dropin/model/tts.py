"""``model.tts`` of the reference layout -> ``dexb200.model.tts`` (see model/__init__.py)."""
from dexb200.model.tts import *  # noqa: F401,F403
from dexb200.model.tts import DeXTTS, GeDEXTTS  # noqa: F401,E402
