"""hevcbitstream_b200 -- B200-native batched bitstream hot path of leslie-wang/hevcbitstream.

The product is the C-ABI library libhevcb200.so (include/hevcb.h, sources in csrc/).  This package is the
thin Python mirror used by tests and bench.py: ctypes bindings plus torch for device memory and streams.
There is no CPU implementation: importing works anywhere, but every compute call needs a B200.
"""
from ._lib import LIB_PATH, load_library, HevcbError  # noqa: F401
from .api import Context, HostIndex, ScanResult  # noqa: F401

__all__ = ["Context", "ScanResult", "HevcbError", "load_library", "LIB_PATH"]
