"""tethys_speech_b200 — B200-native (sm_100a) drop-in for the data-parallel train step of tethys-speech's
Whisper and Wav2Vec2 models. Python host (this package) → ctypes C-ABI (include/tethys.h) → hand-written CUDA
kernels (csrc/). See DESIGN.md for the path, the boundary and the kernels."""
from . import _lib  # noqa: F401
from ._lib import TethysError  # noqa: F401

__all__ = ["_lib", "TethysError"]
