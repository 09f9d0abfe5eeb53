"""``model.diffusion`` of the reference layout -> ``dexb200.model.diffusion`` (see model/__init__.py)."""
from dexb200.model.diffusion import *  # noqa: F401,F403
from dexb200.model.diffusion import Diffusion, GeDiffusion  # noqa: F401,E402
