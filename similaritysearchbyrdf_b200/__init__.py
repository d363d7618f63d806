"""similaritysearchbyrdf_b200 — B200-native (sm_100a) hot path of Dynamic Partition Forest.

The product is libdpf_b200.so (C ABI: include/dpf.h).  This package is the host-side mirror of the reference's
Scala facades (mclab.deploy.{LSHServer, DensevectorRDFInit, SparsevectorRDFInit}, mclab.lsh.LSH) over that ABI.
"""
from . import _lib  # noqa: F401
from .index import DPFIndex  # noqa: F401

__all__ = ["DPFIndex"]
