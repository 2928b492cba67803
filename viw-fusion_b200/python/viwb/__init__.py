"""viwb -- Python plumbing (ctypes) over the B200-native sliding-window backend libviwb.so.

The product is the CUDA library in viw-fusion_b200/csrc behind the C ABI of include/viwb.h; this package
only binds it for tests and bench.py and hosts the seeded synthetic generator.
"""
from . import abi  # noqa: F401
