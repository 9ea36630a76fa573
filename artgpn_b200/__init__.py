"""artgpn_b200 -- B200-native likelihood / prediction path of artgpn.

Drop-in surface (same names as the reference package, artgpn/__init__.py:9-19):
``artgpn_b200.arttwo.network``, ``artgpn_b200.node``, ``artgpn_b200.weight``, ``artgpn_b200.mean``.
All numerical work happens in ``libartgpn_b200.so`` (hand-written sm_100a CUDA, C ABI in
``include/artgpn_b200.h``); there is no CPU fallback.
"""
from . import weight
from . import node
from . import mean
from . import arttwo
from . import art
from .arttwo import network
from ._kernelbase import set_device, get_device
from ._lib import AgpnError, LIB_PATH

__all__ = ["weight", "node", "mean", "arttwo", "art", "network", "set_device", "get_device", "AgpnError", "LIB_PATH"]
__version__ = "0.1.0"
