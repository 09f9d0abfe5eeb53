"""``model.augmentation`` of the reference layout -> ``dexb200.model.augmentation`` (see model/__init__.py)."""
from dexb200.model.augmentation import Augment  # noqa: F401
