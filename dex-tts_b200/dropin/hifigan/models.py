from dexb200.hifigan.models import *  # noqa: F401,F403
from dexb200.hifigan.models import Generator  # noqa: F401
