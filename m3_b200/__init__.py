"""m3_b200 - B200-native M3TSZ batch codec (drop-in for m3db/m3's
src/dbnode/encoding/m3tsz hot path).  See DESIGN.md.

Layout: csrc/ (sm_100a kernels + C ABI -> libm3tsz_b200.so), capi.py (ctypes
binding), codec.py (device-resident batch API over torch tensors), encoding.py
(mirror of the reference's Encoder / ReaderIterator / Decoder interface),
sharded.py (multi-GPU partitioning + NCCL all-gather).
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
