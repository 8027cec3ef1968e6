"""Affine EC arithmetic with the reference's exact assertion semantics
(math_utils.py:50-100).  TEST INFRASTRUCTURE -- see oracle/__init__.py."""
from .params import ALPHA, FIELD_PRIME


def div_mod(n, m, p):
    # math_utils.py:50-56 (igcdex there; pow(,-1,) is the same number)
    return n * pow(m % p, -1, p) % p


def ec_add(p1, p2, p=FIELD_PRIME):
    # math_utils.py:59-68
    assert (p1[0] - p2[0]) % p != 0
    m = div_mod(p1[1] - p2[1], p1[0] - p2[0], p)
    x = (m * m - p1[0] - p2[0]) % p
    y = (m * (p1[0] - x) - p1[1]) % p
    return x, y


def ec_neg(pt, p=FIELD_PRIME):
    return pt[0], (-pt[1]) % p


def ec_double(pt, alpha=ALPHA, p=FIELD_PRIME):
    # math_utils.py:79-88
    assert pt[1] % p != 0
    m = div_mod(3 * pt[0] * pt[0] + alpha, 2 * pt[1], p)
    x = (m * m - 2 * pt[0]) % p
    y = (m * (pt[0] - x) - pt[1]) % p
    return x, y


def ec_mult(m, pt, alpha=ALPHA, p=FIELD_PRIME):
    # math_utils.py:91-100, unrolled into a loop with the same add/double order:
    # the recursion peels the low bit first, so the result is built MSB-first.
    assert m > 0
    bits = bin(m)[3:]          # below the leading 1
    # ec_mult(m, P): m even -> ec_mult(m/2, 2P); m odd -> ec_mult(m-1, P) + P.
    # Iteratively: walk from the LSB doubling the base, remembering pending adds.
    pending = []
    base = pt
    while m != 1:
        if m % 2 == 0:
            base = ec_double(base, alpha, p)
            m //= 2
        else:
            pending.append(base)
            m -= 1
    acc = base
    for q in reversed(pending):
        acc = ec_add(acc, q, p)
    return acc
