"""Field / curve constants (reference: signature.py:38-68, pedersen_params.json,
nothing_up_my_sleeve_gen.py:35-91).  TEST INFRASTRUCTURE -- see oracle/__init__.py."""
import json
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_P = json.load(open(os.path.join(_HERE, "..", "stark_perpetual_b200", "data", "curve_params.json")))

FIELD_PRIME = int(_P["FIELD_PRIME"], 16)
FIELD_GEN = _P["FIELD_GEN"]
EC_ORDER = int(_P["EC_ORDER"], 16)
ALPHA = _P["ALPHA"]
BETA = int(_P["BETA"], 16)
N_ELEMENT_BITS_ECDSA = 251   # floor(log2 p), signature.py:47
N_ELEMENT_BITS_HASH = 252    # p.bit_length(), signature.py:50

assert FIELD_PRIME == 2**251 + 17 * 2**192 + 1


def _dbl(pt):
    # affine doubling, math_utils.py:79-88
    x, y = pt
    m = (3 * x * x + ALPHA) * pow(2 * y, -1, FIELD_PRIME) % FIELD_PRIME
    nx = (m * m - 2 * x) % FIELD_PRIME
    return nx, (m * (x - nx) - y) % FIELD_PRIME


def _expand():
    """506-entry table: shift, G, then doubling chains of P0..P3 (lengths 248,4,248,4);
    nothing_up_my_sleeve_gen.py:85-90."""
    b = {k: (int(v[0], 16), int(v[1], 16)) for k, v in _P["BASE_POINTS"].items()}
    pts = [b["SHIFT_POINT"], b["EC_GEN"]]
    for name, n in zip(("P0", "P1", "P2", "P3"), _P["CHAIN_LENGTHS"]):
        q = b[name]
        for _ in range(n):
            pts.append(q)
            q = _dbl(q)
    return pts


CONSTANT_POINTS = _expand()
SHIFT_POINT = CONSTANT_POINTS[0]
MINUS_SHIFT_POINT = (SHIFT_POINT[0], FIELD_PRIME - SHIFT_POINT[1])
EC_GEN = CONSTANT_POINTS[1]

# Montgomery constants, R = 2^256
R = 2**256
R_MOD_P = R % FIELD_PRIME
R2_MOD_P = R * R % FIELD_PRIME
TWO_ADICITY = 192


def root_of_unity(log_n):
    """omega_{2^log_n} = 3^((p-1)/2^log_n)."""
    assert 0 <= log_n <= TWO_ADICITY
    return pow(FIELD_GEN, (FIELD_PRIME - 1) >> log_n, FIELD_PRIME)
