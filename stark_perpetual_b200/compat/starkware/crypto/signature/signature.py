"""GPU-backed mirror of the reference's `starkware.crypto.signature.signature`
(src/starkware/crypto/signature/signature.py): same names, same argument meaning, same exceptions.

pedersen_hash (:296), verify (:217), private_to_stark_key (:109), get_y_coordinate-dependent x-only keys and
sign (:137; the nonce derivation is host-side RFC 6979 as in :117-134, the curve multiplication k*G runs on
the GPU) call libspg through stark_perpetual_b200; batched variants (`*_batch`) are what a service would use.
There is no CPU fallback: without a CUDA device the first call raises SpgError.
"""
import hashlib
import hmac
import json
import math
import os
from typing import Optional, Tuple, Union

import numpy as np

import stark_perpetual_b200 as _spg
from stark_perpetual_b200._lib import ints_to_limbs, limbs_to_ints

_P = json.load(open(os.path.join(os.path.dirname(_spg.__file__), "data", "curve_params.json")))
FIELD_PRIME = int(_P["FIELD_PRIME"], 16)
FIELD_GEN = _P["FIELD_GEN"]
ALPHA = _P["ALPHA"]
BETA = int(_P["BETA"], 16)
EC_ORDER = int(_P["EC_ORDER"], 16)
N_ELEMENT_BITS_ECDSA = math.floor(math.log(FIELD_PRIME, 2))
assert N_ELEMENT_BITS_ECDSA == 251
N_ELEMENT_BITS_HASH = FIELD_PRIME.bit_length()
assert N_ELEMENT_BITS_HASH == 252
SHIFT_POINT = tuple(int(v, 16) for v in _P["BASE_POINTS"]["SHIFT_POINT"])
MINUS_SHIFT_POINT = (SHIFT_POINT[0], FIELD_PRIME - SHIFT_POINT[1])
EC_GEN = tuple(int(v, 16) for v in _P["BASE_POINTS"]["EC_GEN"])

ECPoint = Tuple[int, int]
ECSignature = Tuple[int, int]


class InvalidPublicKeyError(Exception):
    def __init__(self):
        super().__init__("Given x coordinate does not represent any point on the elliptic curve.")


def _ctx():
    return _spg.get_context(0)


# ---------------------------------------------------------------------------------- Pedersen hash
def pedersen_hash_batch(xs, ys):
    """[pedersen_hash(x, y) for x, y in zip(xs, ys)] in one launch."""
    for v in list(xs) + list(ys):
        assert 0 <= v < FIELD_PRIME                                  # signature.py:307
    out, st = _ctx().pedersen_hash2(ints_to_limbs(xs), ints_to_limbs(ys))
    assert not (st == 1).any()
    assert not (st == 2).any(), "Unhashable input."                  # signature.py:313
    return limbs_to_ints(out)


def pedersen_hash(*elements: int) -> int:
    # signature.py:296-318; the shipped table supports at most two elements (:308-310)
    assert len(elements) <= 2, "pedersen_params.json holds constant points for two elements only"
    for x in elements:
        assert 0 <= x < FIELD_PRIME
    if len(elements) == 0:
        return SHIFT_POINT[0]
    out, st = _ctx().pedersen_chain(ints_to_limbs(list(elements)), len(elements))
    assert st[0] != 1
    assert st[0] != 2, "Unhashable input."
    return limbs_to_ints(out)[0]


def pedersen_hash_as_point(*elements: int) -> ECPoint:
    # signature.py:300-318
    assert 1 <= len(elements) <= 2, "pedersen_params.json holds constant points for two elements only"
    for x in elements:
        assert 0 <= x < FIELD_PRIME
    px, py, st = _ctx().pedersen_hash_point(ints_to_limbs(list(elements)), len(elements))
    assert st[0] != 1
    assert st[0] != 2, "Unhashable input."
    return limbs_to_ints(px)[0], limbs_to_ints(py)[0]


# ---------------------------------------------------------------------------------- keys
def get_y_coordinate(stark_key_x_coordinate: int) -> int:
    # signature.py:84-96 (Python ints reduce mod p silently; so does this wrapper)
    y, st = _ctx().get_y_coordinate(ints_to_limbs([stark_key_x_coordinate % FIELD_PRIME]))
    if st[0] == 1:
        raise InvalidPublicKeyError()
    return limbs_to_ints(y)[0]


def get_random_private_key() -> int:
    import secrets
    return secrets.randbelow(EC_ORDER - 1) + 1                        # signature.py:99-101


def is_point_on_curve(x: int, y: int) -> bool:
    return pow(y, 2, FIELD_PRIME) == (pow(x, 3, FIELD_PRIME) + ALPHA * x + BETA) % FIELD_PRIME   # signature.py:193-194


def is_valid_stark_private_key(private_key: int) -> bool:
    return 0 < private_key < EC_ORDER                                 # signature.py:197-201


def is_valid_stark_key(stark_key: int) -> bool:
    # signature.py:204-214
    try:
        get_y_coordinate(stark_key_x_coordinate=stark_key)
    except InvalidPublicKeyError:
        return False
    return True


def mimic_ec_mult_air(m: int, point: ECPoint, shift_point: ECPoint) -> ECPoint:
    # signature.py:176-190; every assertion of the reference surfaces as AssertionError
    assert 0 < m < 2**N_ELEMENT_BITS_ECDSA
    out, st = _ctx().mimic_ec_mult_air(ints_to_limbs([m]),
                                       ints_to_limbs([point[0] % FIELD_PRIME, point[1] % FIELD_PRIME]).reshape(1, 8),
                                       ints_to_limbs([shift_point[0] % FIELD_PRIME, shift_point[1] % FIELD_PRIME]).reshape(1, 8))
    assert st[0] == 0
    o = limbs_to_ints(out.reshape(2, 4))
    return o[0], o[1]


def private_key_to_ec_point_on_stark_curve(priv_key: int) -> ECPoint:
    assert 0 < priv_key < EC_ORDER                                    # signature.py:105
    x, y, st = _ctx().private_to_stark_key(ints_to_limbs([priv_key]), want_y=True)
    assert st[0] == 0
    return limbs_to_ints(x)[0], limbs_to_ints(y)[0]


def private_to_stark_key(priv_key: int) -> int:
    assert 0 < priv_key < EC_ORDER
    out, st = _ctx().private_to_stark_key(ints_to_limbs([priv_key]))
    assert st[0] == 0
    return limbs_to_ints(out)[0]


def private_to_stark_key_batch(priv_keys):
    for k in priv_keys:
        assert 0 < k < EC_ORDER
    out, st = _ctx().private_to_stark_key(ints_to_limbs(priv_keys))
    assert not st.any()
    return limbs_to_ints(out)


def inv_mod_curve_size(x: int) -> int:
    return pow(x, -1, EC_ORDER)                                       # signature.py:113-114 (div_mod(1, x, n))


# ---------------------------------------------------------------------------------- verify
def verify_batch(msg_hashes, rs, ss, public_keys):
    """public_keys: per element an integer (x-only key, signature.py:229-238) or an (x, y) pair (:239-241); the two kinds
    may be mixed.  Returns a list of bool; raises AssertionError if the reference would raise for any element
    (signature.py:219, :225-227, :241).  Key coordinates are reduced mod p before they go to the device: the reference's
    arithmetic does the same implicitly (get_y_coordinate :90, is_point_on_curve :200 and every EC formula work mod p)."""
    import numbers
    n = len(msg_hashes)
    assert len(rs) == n and len(ss) == n and len(public_keys) == n
    if n == 0:
        return []
    for v in list(msg_hashes) + list(rs) + list(ss):
        assert 0 <= v < 2**256, "operand does not fit 256 bits"
    is_x = [isinstance(k, numbers.Integral) for k in public_keys]
    out = [None] * n
    for kind in (True, False):
        sel = [i for i in range(n) if is_x[i] == kind]
        if not sel:
            continue
        if kind:
            keys_x, keys_y = [int(public_keys[i]) % FIELD_PRIME for i in sel], None
        else:
            for i in sel:
                assert len(public_keys[i]) == 2, "public key must be an integer or an (x, y) pair"
            keys_x = [int(public_keys[i][0]) % FIELD_PRIME for i in sel]
            keys_y = [int(public_keys[i][1]) % FIELD_PRIME for i in sel]
        st = _ctx().ecdsa_verify(ints_to_limbs([int(msg_hashes[i]) for i in sel]), ints_to_limbs([int(rs[i]) for i in sel]),
                                 ints_to_limbs([int(ss[i]) for i in sel]), ints_to_limbs(keys_x),
                                 ints_to_limbs(keys_y) if keys_y is not None else None)
        assert not (st == 2).any(), "precondition violated (s, r, w or msg_hash out of range, or key not on curve)"
        for i, v in zip(sel, st):
            out[i] = bool(v)
    return out


def verify(msg_hash: int, r: int, s: int, public_key: Union[int, ECPoint]) -> bool:
    # signature.py:217-260 -- range assertions first, in the reference's order, with its messages
    assert 1 <= s < EC_ORDER, "s = %s" % s
    w = inv_mod_curve_size(s)
    assert 1 <= r < 2**N_ELEMENT_BITS_ECDSA, "r = %s" % r
    assert 1 <= w < 2**N_ELEMENT_BITS_ECDSA, "w = %s" % w
    assert 0 <= msg_hash < 2**N_ELEMENT_BITS_ECDSA, "msg_hash = %s" % msg_hash
    return verify_batch([msg_hash], [r], [s], [public_key])[0]


# ---------------------------------------------------------------------------------- sign
def _bits2int(data, qlen):
    x = int.from_bytes(data, "big")
    l = len(data) * 8
    return x >> (l - qlen) if l > qlen else x


def _rfc6979_generate_k(order, secexp, hash_func, data, extra_entropy=b""):
    """RFC 6979 section 3.2 (what the reference gets from ecdsa.rfc6979.generate_k, signature.py:25,128)."""
    qlen = order.bit_length()
    holen = hash_func().digest_size
    rolen = (qlen + 7) // 8
    z1 = _bits2int(data, qlen)
    z2 = z1 - order if z1 >= order else z1
    bx = secexp.to_bytes(rolen, "big") + z2.to_bytes(rolen, "big") + extra_entropy
    v, k = b"\x01" * holen, b"\x00" * holen
    k = hmac.new(k, v + b"\x00" + bx, hash_func).digest()
    v = hmac.new(k, v, hash_func).digest()
    k = hmac.new(k, v + b"\x01" + bx, hash_func).digest()
    v = hmac.new(k, v, hash_func).digest()
    while True:
        t = b""
        while len(t) < rolen:
            v = hmac.new(k, v, hash_func).digest()
            t += v
        secret = _bits2int(t, qlen)
        if 1 <= secret < order:
            return secret
        k = hmac.new(k, v + b"\x00", hash_func).digest()
        v = hmac.new(k, v, hash_func).digest()


def generate_k_rfc6979(msg_hash: int, priv_key: int, seed: Optional[int] = None) -> int:
    # signature.py:117-134
    if 1 <= msg_hash.bit_length() % 8 <= 4 and msg_hash.bit_length() >= 248:
        msg_hash *= 16
    extra = b"" if seed is None else seed.to_bytes(math.ceil(seed.bit_length() / 8), "big")
    return _rfc6979_generate_k(EC_ORDER, priv_key, hashlib.sha256,
                               msg_hash.to_bytes(math.ceil(msg_hash.bit_length() / 8), "big"), extra_entropy=extra)


_SEED_LIMIT = 2**64 - 256      # room for the retry counter in the kernel's 64-bit seed


def sign(msg_hash: int, priv_key: int, seed: Optional[int] = None) -> ECSignature:
    # signature.py:137-173
    assert 0 <= msg_hash < 2**N_ELEMENT_BITS_ECDSA, "Message not signable."
    if 1 <= priv_key < EC_ORDER and (seed is None or 0 <= seed < _SEED_LIMIT):
        return sign_batch([msg_hash], [priv_key], [seed])[0]
    # keys outside [1, n) and seeds beyond 64 bits are outside the batch kernel's domain: the reference's loop with
    # the nonce derived on the host and x(k*G) on the GPU
    while True:
        k = generate_k_rfc6979(msg_hash, priv_key, seed)
        seed = 1 if seed is None else seed + 1
        r = private_to_stark_key(k)
        if not (1 <= r < 2**N_ELEMENT_BITS_ECDSA):
            continue
        if (msg_hash + r * priv_key) % EC_ORDER == 0:
            continue
        w = k * pow(msg_hash + r * priv_key, -1, EC_ORDER) % EC_ORDER
        if not (1 <= w < 2**N_ELEMENT_BITS_ECDSA):
            continue
        return r, inv_mod_curve_size(w)


def sign_batch(msg_hashes, priv_keys, seeds=None):
    """[sign(m, k, seed) for ...] in one launch (spg_sign_batch): each GPU thread derives its RFC 6979 nonce
    (HMAC-SHA256, signature.py:117-134), multiplies the generator, finishes the signature modulo n and retries with
    the bumped seed on the reference's three rejection rules (signature.py:158-170).  Keys must lie in [1, n)."""
    n = len(msg_hashes)
    for m in msg_hashes:
        assert 0 <= m < 2**N_ELEMENT_BITS_ECDSA, "Message not signable."
    for k in priv_keys:
        assert 1 <= k < EC_ORDER, "private key outside [1, EC_ORDER)"
    sd = None
    if seeds is not None:
        sd = [0 if x is None else x for x in seeds]      # None and 0 are the same signature (no extra entropy)
        for x in sd:
            assert 0 <= x < _SEED_LIMIT
        sd = np.array(sd, dtype=np.uint64)
    if n == 0:
        return []
    r, s, st = _ctx().sign(ints_to_limbs(msg_hashes), ints_to_limbs(priv_keys), sd)
    assert not st.any(), "spg_sign_batch status %s" % sorted(set(st.tolist()))
    return list(zip(limbs_to_ints(r), limbs_to_ints(s)))


def pedersen_merkle_root(leaves):
    """Root of the Merkle tree with pedersen_hash nodes over a power-of-two number of leaves (the StarkEx state-tree
    node function, src/services/perpetual/cairo/state/state.cairo:155-173), computed level by level on the GPU."""
    for v in leaves:
        assert 0 <= v < FIELD_PRIME
    root, _nodes, st = _ctx().pedersen_merkle_tree(ints_to_limbs(leaves))
    assert st != 1
    assert st != 2, "Unhashable input."
    return limbs_to_ints(root.reshape(1, 4))[0]


def grind_key(key_seed: int, key_value_limit: int) -> int:
    # signature.py:263-288 (host-side hashing only)
    max_allowed = 2**256 - (2**256 % key_value_limit)

    def nb(x):
        return x.to_bytes(max(1, -(-x.bit_length() // 8)), "big")
    index = 0
    while True:
        key = int(hashlib.sha256(nb(key_seed) + nb(index)).hexdigest(), 16)
        if key < max_allowed:
            return key % key_value_limit
        index += 1
