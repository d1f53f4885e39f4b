"""B200-native GIM adversarial-training hot path (drop-in for the reference's models/ + training/ trainers).

Host side: the reference's nn.Module / trainer / checkpoint surface.  Device side: libgim_b200.so (hand-written sm_100a
CUDA behind the C ABI in include/gim_b200.h).  No CPU fallback: importing works anywhere, computing needs the library and
a CUDA device.
"""
from . import ops
from .ops import set_precision, get_precision, set_conv_algo, set_deterministic

__all__ = ["ops", "set_precision", "get_precision", "set_conv_algo", "set_deterministic"]
