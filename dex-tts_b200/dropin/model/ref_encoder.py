"""``model.ref_encoder`` of the reference layout -> ``dexb200.model.ref_encoder`` (see model/__init__.py)."""
from dexb200.model.ref_encoder import *  # noqa: F401,F403
from dexb200.model.ref_encoder import LF0Encoder, TIVEncoder, TVEncoder, style_fusion  # noqa: F401,E402
