"""Host side of the second AIR: a batch of ECDSA verifications proven as one trace (csrc/air_ecdsa.cu, csrc/prove.cu).

Each 256-row block of the trace is one `verify(msg_hash, r, s, public_key)` of the reference
(src/starkware/crypto/signature/signature.py:217-260): `w = inv_mod_curve_size(s)` is computed here exactly as line 219
does, the three `mimic_ec_mult_air` walks and the final `r == x` comparison are the trace.  The constraint system is this
repo's own; the proof is checked by oracle/stark.py `verify` (header VERSION 2) in the tests.
"""
import json
import os

from ._lib import get_context, ints_to_limbs

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "curve_params.json")) as _f:
    EC_ORDER = int(json.load(_f)["EC_ORDER"], 16)                     # signature.py:43

BLOCK = 256


def air_inputs(msgs, r, s, keys):
    """msgs, r, s: ints; keys: (x, y) curve points.  -> the five (n, 4) uint64 arrays spg_ecdsa_air_trace takes.
    ValueError where verify's own range checks fail (signature.py:219-227)."""
    w = []
    for si in s:
        if not 1 <= si < EC_ORDER:
            raise ValueError("s out of range")
        w.append(pow(si, -1, EC_ORDER))
    return (ints_to_limbs(msgs), ints_to_limbs(r), ints_to_limbs(w), ints_to_limbs([k[0] for k in keys]),
            ints_to_limbs([k[1] for k in keys]))


def prove_signatures(msgs, r, s, keys, n_queries=30, ctx=None):
    """Proof that all len(msgs) = 2^k >= 2 signatures verify; the statement (public input, carried by the proof) is the
    list of (msg_hash, key x).  Returns (proof bytes, log_n)."""
    ctx = ctx or get_context()
    n = len(msgs)
    if n < 2 or n & (n - 1):
        raise ValueError("the batch must hold a power of two (>= 2) signatures")
    log_n = (n * BLOCK).bit_length() - 1
    m_, r_, w_, kx, ky = air_inputs(msgs, r, s, keys)
    trace = ctx.ecdsa_air_trace(log_n, m_, r_, w_, kx, ky)
    return ctx.prove_ecdsa(trace, log_n, m_, kx, n_queries), log_n
