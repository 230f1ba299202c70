"""avddpg_b200 -- B200-native (sm_100a) implementation of the avddpg data-parallel hot path:
batched vehicle-platoon environment + OU noise + GPU replay + DDPG learn step + federated aggregation.

Host code is Python with PyTorch as a thin FFI (device memory, streams, torch.distributed); all
arithmetic runs in hand-written CUDA kernels behind the C ABI in include/avddpg_b200.h
(avddpg_b200/lib/libavddpg_b200.so).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .config import Config  # noqa: F401

__all__ = ["Config", "environment", "noise", "replaybuffer"]
