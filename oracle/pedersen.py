"""Pedersen hash (signature.py:296-318; fast variant fast_pedersen_hash.py:26-44).
TEST INFRASTRUCTURE -- see oracle/__init__.py."""
from .curve import ec_add
from .params import CONSTANT_POINTS, FIELD_PRIME, N_ELEMENT_BITS_HASH, SHIFT_POINT


def pedersen_hash_as_point(*elements, trace=None):
    """signature.py:300-318.  `trace`, if a list, receives every partial sum (row by row) --
    this is the Pedersen-builtin AIR witness."""
    point = SHIFT_POINT
    for i, x in enumerate(elements):
        assert 0 <= x < FIELD_PRIME
        pts = CONSTANT_POINTS[2 + i * N_ELEMENT_BITS_HASH:2 + (i + 1) * N_ELEMENT_BITS_HASH]
        assert len(pts) == N_ELEMENT_BITS_HASH
        for pt in pts:
            if trace is not None:
                trace.append((point, x))
            assert point[0] != pt[0], "Unhashable input."
            if x & 1:
                point = ec_add(point, pt, FIELD_PRIME)
            x >>= 1
        assert x == 0
    return point


def pedersen_hash(*elements):
    return pedersen_hash_as_point(*elements)[0]


def pedersen_hash_func(x: bytes, y: bytes) -> bytes:
    # fast_pedersen_hash.py:47-52: 32-byte big-endian in and out
    assert len(x) == len(y) == 32, "Unexpected element length."
    return pedersen_hash(int.from_bytes(x, "big"), int.from_bytes(y, "big")).to_bytes(32, "big")
