"""vstnet_b200 — B200-native (sm_100a) CAP-VSTNet stylization hot path.

Public surface mirrors the reference's ``models`` package for this path:
``RevResNet`` (encode / decode) and ``cWCT`` (whitening-colouring transform).
"""
from .RevResNet import RevResNet  # noqa: F401
from .cWCT import cWCT  # noqa: F401

__all__ = ["RevResNet", "cWCT"]
__version__ = "0.1.0"
