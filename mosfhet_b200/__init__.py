"""mosfhet_b200 -- B200 (sm_100a) implementation of MOSFHET's programmable-bootstrap hot path.

The product is ``libmosfhet_b200.so`` (hand-written CUDA behind a C ABI, ``include/mosfhet_b200.h``);
this package is its Python host side: ctypes handle types mirroring the reference (``abi``), the
same operator names over those handles and flat/device-resident batch calls (``api``), ciphertext
sharding across GPUs (``sharding``), and the in-tree build (``build``).
"""
from . import abi  # noqa: F401

__all__ = ["abi", "api", "build", "params", "sharding"]
