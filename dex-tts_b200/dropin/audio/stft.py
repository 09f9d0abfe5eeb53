from dexb200.audio.stft import *  # noqa: F401,F403
from dexb200.audio.stft import TacotronSTFT, slaney_mel_basis  # noqa: F401,E402
