"""CPU: Python NTT oracle self-consistency and the C oracle against it."""
import random

import numpy as np

from oracle import clib, ntt as ontt
from oracle.params import FIELD_PRIME as P, root_of_unity
from stark_perpetual_b200._lib import ints_to_limbs, limbs_to_ints


def test_python_ntt_matches_definition():
    rng = random.Random(5)
    n = 16
    a = [rng.randrange(P) for _ in range(n)]
    w = root_of_unity(4)
    direct = [sum(a[i] * pow(w, i * k, P) for i in range(n)) % P for k in range(n)]
    assert ontt.ntt(a) == direct
    assert ontt.ntt(ontt.ntt(a), inverse=True) == a


def test_lde_agrees_with_horner():
    rng = random.Random(6)
    n, lb = 8, 2
    col = [rng.randrange(P) for _ in range(n)]
    coeffs = ontt.ntt(col, inverse=True)
    cosets = ontt.lde(col, lb)
    wb, wn = root_of_unity(5), root_of_unity(3)
    for j in range(4):
        for i in range(n):
            x = 3 * pow(wb, j, P) * pow(wn, i, P) % P
            assert cosets[j][i] == sum(c * pow(x, k, P) for k, c in enumerate(coeffs)) % P


def test_c_oracle_field_and_ntt():
    rng = random.Random(7)
    a = [rng.randrange(P) for _ in range(64)] + [0, 1, P - 1]
    b = [rng.randrange(P) for _ in range(64)] + [P - 1, P - 1, P - 1]
    got = limbs_to_ints(clib.mul_batch(ints_to_limbs(a), ints_to_limbs(b)))
    assert got == [x * y % P for x, y in zip(a, b)]
    for log_n in (0, 1, 5, 10):
        n = 1 << log_n
        v = [rng.randrange(P) for _ in range(2 * n)]
        arr = ints_to_limbs(v)
        for inverse in (False, True):
            want = ontt.ntt(v[:n], inverse) + ontt.ntt(v[n:], inverse)
            assert limbs_to_ints(clib.ntt(arr, log_n, inverse, 2)) == want
            rev = ontt.bitrev_permute(want[:n]) + ontt.bitrev_permute(want[n:])
            assert limbs_to_ints(clib.ntt(arr, log_n, inverse, 0)) == rev
            vin = ontt.bitrev_permute(v[:n]) + ontt.bitrev_permute(v[n:])
            assert limbs_to_ints(clib.ntt(ints_to_limbs(vin), log_n, inverse, 1)) == want
