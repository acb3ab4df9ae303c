"""pcdb200 — Python mirror of the C-ABI in include/pcdb200.h (tests and bench drive the CUDA library through it).

The product is the shared library point-cloud-donkey_b200/csrc/libpcdb200.so; this package is plumbing only.
"""
from .structs import *  # noqa: F401,F403
