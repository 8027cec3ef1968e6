"""GPU-backed mirror of the reference's `starkware.crypto.signature.fast_pedersen_hash`
(src/starkware/crypto/signature/fast_pedersen_hash.py:34-52): the same two entry points; the 32-byte big-endian byte
ABI of `pedersen_hash_func` goes straight to spg_pedersen_hash2_batch_be32."""
import numpy as np

import stark_perpetual_b200 as _spg
from starkware.crypto.signature.signature import FIELD_PRIME, pedersen_hash as _hash

HASH_BYTES = 32


def pedersen_hash(x: int, y: int) -> int:
    # fast_pedersen_hash.py:34-44 (same value as signature.pedersen_hash)
    assert 0 <= x < FIELD_PRIME and 0 <= y < FIELD_PRIME
    return _hash(x, y)


def pedersen_hash_func(x: bytes, y: bytes) -> bytes:
    # fast_pedersen_hash.py:47-52
    assert len(x) == HASH_BYTES and len(y) == HASH_BYTES
    out, st = _spg.get_context(0).pedersen_hash2_be32(np.frombuffer(x, dtype=np.uint8), np.frombuffer(y, dtype=np.uint8))
    assert st[0] != 1
    assert st[0] != 2, "Unhashable input."
    return out[0].tobytes()


def pedersen_hash_func_batch(xs: bytes, ys: bytes) -> bytes:
    """n concatenated 32-byte inputs each -> n concatenated 32-byte digests, one launch."""
    assert len(xs) == len(ys) and len(xs) % HASH_BYTES == 0
    out, st = _spg.get_context(0).pedersen_hash2_be32(np.frombuffer(xs, dtype=np.uint8), np.frombuffer(ys, dtype=np.uint8))
    assert not (st == 1).any()
    assert not (st == 2).any(), "Unhashable input."
    return out.tobytes()
