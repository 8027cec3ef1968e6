"""stark_perpetual_b200 -- B200-native (sm_100a CUDA) implementation of the data-parallel hot path
behind StarkEx Perpetual: field/NTT/LDE/AIR/FRI/Merkle prover stages and batched Pedersen / ECDSA.

The compute lives in libspg.so (C-ABI, include/spg.h); this package is the thin ctypes host
side that mirrors the reference's Python interface.  There is no CPU fallback: importing works
anywhere, but creating a Context without a CUDA device raises.
"""
from ._lib import Context, SpgError, get_context, lib_path  # noqa: F401

__all__ = ["Context", "SpgError", "get_context", "lib_path"]
