"""Drop-in import path of the reference's ``models/segmentation/SegReMapping.py`` (device implementation)."""
from vstnet_b200.segmentation import SegReMapping  # noqa: F401
