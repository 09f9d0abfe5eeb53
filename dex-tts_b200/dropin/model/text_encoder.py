"""``model.text_encoder`` of the reference layout -> ``dexb200.model.text_encoder`` (see model/__init__.py)."""
from dexb200.model.text_encoder import *  # noqa: F401,F403
from dexb200.model.text_encoder import GeTextEncoder, TextEncoder  # noqa: F401,E402
