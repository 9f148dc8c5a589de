"""python_5gtoolbox_b200 -- B200-native (sm_100a) batched 5G NR LDPC engine.

A drop-in for the LDPC hot path of xu753x/python_5gtoolbox (``py5gphy/ldpc``): the same call
signatures (``python_5gtoolbox_b200.ldpc.nr_ldpc_encode.encode_ldpc`` ...), computed by hand-written
CUDA kernels behind the C ABI of ``include/nrldpc_b200.h``.  There is no CPU fallback: every compute
call raises if ``libnrldpc_b200.so`` or a CUDA device is missing.

Batched entry points live in :mod:`python_5gtoolbox_b200.engine`; :func:`install` overlays the
drop-in onto an importable reference ``py5gphy`` so its scripts run unchanged.
"""
from ._lib import NrLdpcError, lib, build  # noqa: F401
from .overlay import install, uninstall  # noqa: F401

__version__ = "0.1.0"
