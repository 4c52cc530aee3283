"""sefd: B200-native (sm_100a) DCCRN speech-enhancement train-step path behind a C ABI (libsefd.so)."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
