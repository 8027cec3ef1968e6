"""STARK-curve ECDSA (signature.py:84-260) incl. RFC 6979 nonce generation
(the `ecdsa` package's rfc6979.generate_k, pinned ecdsa==0.17.0, is not in the image;
restated here from RFC 6979 section 3.2 + the retry rule of that package).
TEST INFRASTRUCTURE -- see oracle/__init__.py."""
import hashlib
import hmac
import math

from .curve import div_mod, ec_add, ec_double, ec_mult
from .params import (ALPHA, BETA, EC_GEN, EC_ORDER, FIELD_PRIME, MINUS_SHIFT_POINT,
                     N_ELEMENT_BITS_ECDSA, SHIFT_POINT, TWO_ADICITY)


class InvalidPublicKeyError(Exception):
    def __init__(self):
        super().__init__("Given x coordinate does not represent any point on the elliptic curve.")


def is_quad_residue(a, p=FIELD_PRIME):
    a %= p
    return a < 2 or pow(a, (p - 1) // 2, p) == 1     # sympy: 0 and 1 are residues


def sqrt_mod(a, p=FIELD_PRIME):
    """min of the two roots (math_utils.py:43-47); Tonelli-Shanks."""
    a %= p
    if a == 0:
        return 0
    q, s = (p - 1) >> TWO_ADICITY, TWO_ADICITY
    c, x, t, m = pow(3, q, p), pow(a, (q + 1) // 2, p), pow(a, q, p), s
    while t != 1:
        i, tt = 0, t
        while tt != 1:
            tt = tt * tt % p
            i += 1
        b = pow(c, 1 << (m - i - 1), p)
        x, c = x * b % p, b * b % p
        t, m = t * c % p, i
    return min(x, p - x)


def get_y_coordinate(x):
    # signature.py:84-96
    y2 = (x * x * x + ALPHA * x + BETA) % FIELD_PRIME
    if not is_quad_residue(y2):
        raise InvalidPublicKeyError()
    return sqrt_mod(y2)


def private_key_to_ec_point_on_stark_curve(priv):
    assert 0 < priv < EC_ORDER
    return ec_mult(priv, EC_GEN, ALPHA, FIELD_PRIME)


def private_to_stark_key(priv):
    return private_key_to_ec_point_on_stark_curve(priv)[0]


def inv_mod_curve_size(x):
    return div_mod(1, x, EC_ORDER)


# ---- RFC 6979 (HMAC-SHA256 DRBG), as python-ecdsa 0.17 implements it -------------------------
def _bits2int(data, qlen):
    x = int.from_bytes(data, "big")
    l = len(data) * 8
    return x >> (l - qlen) if l > qlen else x


def _bits2octets(data, order):
    z1 = _bits2int(data, order.bit_length())
    z2 = z1 - order
    if z2 < 0:
        z2 = z1
    return z2.to_bytes((order.bit_length() + 7) // 8, "big")


def rfc6979_generate_k(order, secexp, hash_func, data, retry_gen=0, extra_entropy=b""):
    qlen = order.bit_length()
    holen = hash_func().digest_size
    rolen = (qlen + 7) // 8
    bx = secexp.to_bytes(rolen, "big") + _bits2octets(data, order) + extra_entropy
    v = b"\x01" * holen
    k = b"\x00" * holen
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
            if retry_gen <= 0:
                return secret
            retry_gen -= 1
        k = hmac.new(k, v + b"\x00", hash_func).digest()
        v = hmac.new(k, v, hash_func).digest()


def generate_k_rfc6979(msg_hash, priv_key, seed=None):
    # signature.py:117-134
    if 1 <= msg_hash.bit_length() % 8 <= 4 and msg_hash.bit_length() >= 248:
        msg_hash *= 16
    extra = b"" if seed is None else seed.to_bytes(math.ceil(seed.bit_length() / 8), "big")
    return rfc6979_generate_k(
        EC_ORDER, priv_key, hashlib.sha256,
        msg_hash.to_bytes(math.ceil(msg_hash.bit_length() / 8), "big"), extra_entropy=extra)


def sign(msg_hash, priv_key, seed=None):
    # signature.py:137-173
    assert 0 <= msg_hash < 2**N_ELEMENT_BITS_ECDSA, "Message not signable."
    while True:
        k = generate_k_rfc6979(msg_hash, priv_key, seed)
        seed = 1 if seed is None else seed + 1
        x = ec_mult(k, EC_GEN, ALPHA, FIELD_PRIME)[0]
        r = int(x)
        if not (1 <= r < 2**N_ELEMENT_BITS_ECDSA):
            continue
        if (msg_hash + r * priv_key) % EC_ORDER == 0:
            continue
        w = div_mod(k, msg_hash + r * priv_key, EC_ORDER)
        if not (1 <= w < 2**N_ELEMENT_BITS_ECDSA):
            continue
        return r, inv_mod_curve_size(w)


def mimic_ec_mult_air(m, point, shift_point, trace=None):
    # signature.py:176-190
    assert 0 < m < 2**N_ELEMENT_BITS_ECDSA
    partial_sum = shift_point
    for _ in range(N_ELEMENT_BITS_ECDSA):
        if trace is not None:
            trace.append((partial_sum, point, m))
        assert partial_sum[0] != point[0]
        if m & 1:
            partial_sum = ec_add(partial_sum, point, FIELD_PRIME)
        point = ec_double(point, ALPHA, FIELD_PRIME)
        m >>= 1
    assert m == 0
    return partial_sum


def is_point_on_curve(x, y):
    return pow(y, 2, FIELD_PRIME) == (pow(x, 3, FIELD_PRIME) + ALPHA * x + BETA) % FIELD_PRIME


def is_valid_stark_key(x):
    try:
        get_y_coordinate(x)
    except InvalidPublicKeyError:
        return False
    return True


def verify(msg_hash, r, s, public_key):
    # signature.py:217-260
    assert 1 <= s < EC_ORDER, "s = %s" % s
    w = inv_mod_curve_size(s)
    assert 1 <= r < 2**N_ELEMENT_BITS_ECDSA, "r = %s" % r
    assert 1 <= w < 2**N_ELEMENT_BITS_ECDSA, "w = %s" % w
    assert 0 <= msg_hash < 2**N_ELEMENT_BITS_ECDSA, "msg_hash = %s" % msg_hash
    if isinstance(public_key, int):
        try:
            y = get_y_coordinate(public_key)
        except InvalidPublicKeyError:
            return False
        return verify(msg_hash, r, s, (public_key, y)) or verify(
            msg_hash, r, s, (public_key, (-y) % FIELD_PRIME))
    assert is_point_on_curve(x=public_key[0], y=public_key[1])
    try:
        zG = mimic_ec_mult_air(msg_hash, EC_GEN, MINUS_SHIFT_POINT)
        rQ = mimic_ec_mult_air(r, public_key, SHIFT_POINT)
        wB = mimic_ec_mult_air(w, ec_add(zG, rQ, FIELD_PRIME), SHIFT_POINT)
        x = ec_add(wB, MINUS_SHIFT_POINT, FIELD_PRIME)[0]
    except AssertionError:
        return False
    return r == x


def grind_key(key_seed, key_value_limit):
    # signature.py:263-288
    max_allowed = 2**256 - (2**256 % key_value_limit)

    def nb(x):
        return x.to_bytes(max(1, -(-x.bit_length() // 8)), "big")
    index = 0
    while True:
        key = int(hashlib.sha256(nb(key_seed) + nb(index)).hexdigest(), 16)
        if key < max_allowed:
            return key % key_value_limit
        index += 1
