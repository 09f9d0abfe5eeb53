"""``model.utils`` of the reference layout -> ``dexb200.model.utils`` (see model/__init__.py)."""
from dexb200.model.utils import *  # noqa: F401,F403
from dexb200.model.utils import align_durations, fix_len_compatibility, sequence_mask  # noqa: F401,E402
