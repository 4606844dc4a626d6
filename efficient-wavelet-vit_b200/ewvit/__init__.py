"""Python binding of libewvit.so (ctypes over the C ABI in include/ewvit.h).

There is NO fallback: importing :mod:`ewvit.ops` works without a GPU (so that host logic can be
unit-tested), but every compute entry point raises if the library or a B200 is missing.
"""
from ._lib import EwvitError, lib, lib_path, load  # noqa: F401
