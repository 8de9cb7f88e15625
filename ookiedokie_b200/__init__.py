"""ookiedokie_b200 -- B200 (sm_100a) receive path for OOKiedokie.

The product is the C ABI in include/ookd_gpu.h (ookiedokie_b200/lib/libookd_gpu.so, built from
ookiedokie_b200/csrc) and the C host front end in ookiedokie_b200/host.  `binding` is a ctypes view of
that ABI for tests, bench.py and scripting.
"""
from . import binding  # noqa: F401

__all__ = ["binding"]
