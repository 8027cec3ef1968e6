"""Radix-2 NTT / LDE over the STARK prime, plain Python ints.  PARITY UNPINNED (no reference
NTT exists; field and generator from signature.py:41-42).  TEST INFRASTRUCTURE."""
from .params import FIELD_PRIME as P, root_of_unity


def bitrev(i, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def bitrev_permute(a):
    n = len(a)
    bits = n.bit_length() - 1
    return [a[bitrev(i, bits)] for i in range(n)]


def ntt(a, inverse=False):
    """Natural order in, natural order out.  X[k] = sum_i a[i] w^(ik), w = 3^((p-1)/n) (or its
    inverse, with the 1/n scale)."""
    n = len(a)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    w = root_of_unity(log_n)
    if inverse:
        w = pow(w, -1, P)
    a = bitrev_permute(list(a))
    h = 1
    while h < n:
        wh = pow(w, n // (2 * h), P)
        tw = [1] * h
        for j in range(1, h):
            tw[j] = tw[j - 1] * wh % P
        for b in range(0, n, 2 * h):
            for j in range(h):
                u = a[b + j]
                t = a[b + j + h] * tw[j] % P
                a[b + j] = (u + t) % P
                a[b + j + h] = (u - t) % P
        h *= 2
    if inverse:
        ninv = pow(n, -1, P)
        a = [x * ninv % P for x in a]
    return a


def lde(column, log_blowup, offset=3):
    """Evaluations of the degree < N interpolant of `column` (values on <w_N>, natural order) on
    the cosets  offset * w_{BN}^j * <w_N>,  j = 0..B-1, each in natural order.  Returns a list of B
    lists (coset-major layout of DESIGN.md)."""
    n = len(column)
    log_n = n.bit_length() - 1
    coeffs = ntt(column, inverse=True)
    wb = root_of_unity(log_n + log_blowup)
    out = []
    for j in range(1 << log_blowup):
        s = offset * pow(wb, j, P) % P
        sc, acc = [], 1
        for k in range(n):
            sc.append(coeffs[k] * acc % P)
            acc = acc * s % P
        out.append(ntt(sc))
    return out
