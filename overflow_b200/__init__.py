"""overflow_b200 -- B200-native D8 flow routing (flow direction + flow accumulation).

Drop-in for the hot path of Denver-Automation-Analytics/overflow: the public functions keep
the reference's names, arguments and results; the compute is hand-written CUDA for sm_100a
behind the C ABI in include/overflow_b200.h.  No CPU fallback.
"""
from . import constants  # noqa: F401
from ._native import OverflowB200Error  # noqa: F401

__version__ = "0.1.0"
