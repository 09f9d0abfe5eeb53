"""``model.monotonic_align`` of the reference layout -> ``dexb200.model.monotonic_align`` (see model/__init__.py)."""
from dexb200.model.monotonic_align import maximum_path  # noqa: F401
