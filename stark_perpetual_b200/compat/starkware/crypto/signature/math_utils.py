"""Drop-in for src/starkware/crypto/signature/math_utils.py backed by libspg (same names, arguments and exceptions).

The curve operations and the modular helpers run on the GPU for the STARK prime -- the only modulus the reference's hot
path uses them with (signature.py:84-96, 104-110, 155, 176-190).  A different modulus is outside libspg's domain and
raises ValueError here instead of silently computing on the host (there is no CPU fallback), with one exception:
div_mod modulo the curve order, the scalar inversion of signature.py:113-114, is a single host-side pow().
"""
from typing import Tuple

from stark_perpetual_b200._lib import FIELD_PRIME, get_context, ints_to_limbs, limbs_to_ints

ECPoint = Tuple[int, int]
_EC_ORDER = 0x800000000000010FFFFFFFFFFFFFFFFB781126DCAE7B2321E66A241ADC64D2F


def _only_stark_prime(p, what):
    if p != FIELD_PRIME:
        raise ValueError("%s: libspg implements the STARK prime only (got modulus %#x)" % (what, p))


def pi_as_string(digits: int) -> str:
    # math_utils.py:28-33 (constants provenance, nothing_up_my_sleeve_gen.py); no field arithmetic involved
    import mpmath
    mpmath.mp.dps = digits
    return "3" + str(mpmath.mp.pi)[2:]


def is_quad_residue(n: int, p: int) -> bool:
    _only_stark_prime(p, "is_quad_residue")
    _y, st = get_context().field_sqrt(ints_to_limbs([n % p]))
    return st[0] == 0


def sqrt_mod(n: int, p: int) -> int:
    """The minimum positive m with m*m % p == n (math_utils.py:43-47)."""
    _only_stark_prime(p, "sqrt_mod")
    y, st = get_context().field_sqrt(ints_to_limbs([n % p]))
    if st[0] != 0:
        raise ValueError("min() arg is an empty sequence")      # what the reference's min([]) raises
    return limbs_to_ints(y)[0]


def div_mod(n: int, m: int, p: int) -> int:
    """0 <= x < p with (m * x) % p == n (math_utils.py:50-56)."""
    if p == _EC_ORDER:
        assert m % p != 0
        return n * pow(m, -1, p) % p
    _only_stark_prime(p, "div_mod")
    assert m % p != 0                                            # igcdex(m, p) != 1  <=>  p | m
    ctx = get_context()
    inv = ctx.field_op("inv", ints_to_limbs([m % p]))
    return limbs_to_ints(ctx.field_op("mul", ints_to_limbs([n % p]), inv))[0]


def _xy(points, p):
    return ints_to_limbs([c % p for pt in points for c in pt]).reshape(-1, 8)


def _unxy(out):
    v = limbs_to_ints(out.reshape(-1, 4))
    return [(v[2 * i], v[2 * i + 1]) for i in range(len(v) // 2)]


def ec_add_batch(points1, points2, p: int = FIELD_PRIME):
    _only_stark_prime(p, "ec_add")
    out, st = get_context().ec_op(0, _xy(points1, p), _xy(points2, p))
    assert not st.any()                                          # math_utils.py:64
    return _unxy(out)


def ec_add(point1: ECPoint, point2: ECPoint, p: int) -> ECPoint:
    return ec_add_batch([point1], [point2], p)[0]


def ec_neg(point: ECPoint, p: int) -> ECPoint:
    x, y = point
    return (x, (-y) % p)


def ec_double_batch(points, alpha: int = 1, p: int = FIELD_PRIME):
    _only_stark_prime(p, "ec_double")
    if alpha % p != 1:
        raise ValueError("ec_double: libspg implements the STARK curve (alpha = 1) only")
    out, st = get_context().ec_op(1, _xy(points, p))
    assert not st.any()                                          # math_utils.py:84
    return _unxy(out)


def ec_double(point: ECPoint, alpha: int, p: int) -> ECPoint:
    return ec_double_batch([point], alpha, p)[0]


def ec_mult_batch(ms, points, alpha: int = 1, p: int = FIELD_PRIME):
    _only_stark_prime(p, "ec_mult")
    if alpha % p != 1:
        raise ValueError("ec_mult: libspg implements the STARK curve (alpha = 1) only")
    for m in ms:
        if not 0 < m < 2**256:
            raise RecursionError("ec_mult: m must be positive (the reference recurses forever otherwise)")
    out, st = get_context().ec_op(2, _xy(points, p), ints_to_limbs(ms))
    assert not st.any()                                          # an ec_add / ec_double assertion inside the recursion
    return _unxy(out)


def ec_mult(m: int, point: ECPoint, alpha: int, p: int) -> ECPoint:
    return ec_mult_batch([m], [point], alpha, p)[0]
