"""sorrel_b200 -- B200-native batched gridworld engine behind Sorrel's plugin API.

Environment step, egocentric one-hot observation and replay-buffer add/sample of
social-ai-uoft/sorrel, implemented as hand-written sm_100a CUDA kernels
(`sorrel_b200/csrc`, C ABI in `include/sorrel_b200.h`) for tens of thousands to millions of
independent world instances per launch.  There is no CPU fallback: importing the package is
cheap, but every operation needs `sorrel_b200/lib/libsorrel_b200.so` and a CUDA device.
"""
from ._abi import SorrelB200Error  # noqa: F401

__version__ = "0.1.0"
