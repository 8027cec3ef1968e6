"""The ECDSA-builtin AIR: a batch of signature verifications as an execution trace, plain Python ints.  TEST INFRASTRUCTURE.

What the reference pins is the WITNESS SEMANTICS: every 256-row block walks the three `mimic_ec_mult_air` loops of one
`verify` call (signature.py:176-190 and :243-260), step for step, with each of their assertions turned into an inverse cell:

    zG = mimic_ec_mult_air(msg_hash, EC_GEN, MINUS_SHIFT_POINT)              lane A
    rQ = mimic_ec_mult_air(r, public_key, SHIFT_POINT)                        lane B
    wB = mimic_ec_mult_air(w, ec_add(zG, rQ, FIELD_PRIME), SHIFT_POINT)       lane C
    x = ec_add(wB, MINUS_SHIFT_POINT, FIELD_PRIME)[0];  return r == x

PARITY UNPINNED for the constraint system itself: the reference has no AIR (the builtin's constraints
live in Stone, un-vendored).  The constraints below are this repo's own; the protocol is oracle/stark.py's (prove_air /
verify), unchanged: 25 columns, mask {x, x w_N}, composition degree < 4N.

Trace: N = 2^log_n rows = N/256 blocks; row 256 b + t is step t of block b.  Block b carries lanes A and B of signature b
and lane C of signature b-1 (cyclically): lane C's point is zG + rQ, which lanes A and B hold in their partial sums on
the last row of the block before, so the hand-over is a next-row constraint.  Per lane: M (scalar >> t), (PX, PY) partial
sum BEFORE step t, (QX, QY) the doubled point 2^t Q (lane A: periodic columns 2^t G instead), SA / SD slopes of the
addition / doubling of step t, I = 1 / (PX - QX) (the `partial_sum[0] != point[0]` assertion).  T1 carries r of the
block's signature, T2 the r of the signature lane C is finishing; V1, V2 (row 0) invert the scalars (`0 < m`).
Scalars are unpacked into 251 bits (M = 0 on rows 251..255: `assert m == 0` after N_ELEMENT_BITS_ECDSA steps).
Public input: the list of (msg_hash, key x) of every signature -- the two cells the Cairo ECDSA builtin exposes; (r, w) stay
witness.  The lists enter as two polynomials of degree < N/256 interpolating them over the block-start rows
(x^(N/256) = 1), compared with the cells M_A and QX_B on exactly those rows; the verifier evaluates them at the
out-of-domain point itself.
"""
from .params import ALPHA, BETA, EC_GEN, FIELD_PRIME as P, MINUS_SHIFT_POINT, N_ELEMENT_BITS_ECDSA, SHIFT_POINT, root_of_unity
from .curve import ec_add, ec_double
from . import ntt as ontt
from . import stark

N_COLS = 25
BLOCK = 256
NBITS = N_ELEMENT_BITS_ECDSA                      # 251 steps per scalar
(AM, APX, APY, ASA, AI,
 BM, BPX, BPY, BQX, BQY, BSA, BSD, BI,
 CM, CPX, CPY, CQX, CQY, CSA, CSD, CI,
 T1, T2, V1, V2) = range(N_COLS)
N_ALPHA = 52
KIND = 2
assert ALPHA == 1

inv, ser = stark.inv, stark.ser


def doubled_generator():
    """2^t G for t < 251 (lane A's periodic columns), (0, 0) on the padding rows"""
    gx, gy = [0] * BLOCK, [0] * BLOCK
    q = EC_GEN
    for t in range(NBITS):
        gx[t], gy[t] = q
        q = ec_double(q)
    return gx, gy


def _walk(cols, base, m, point, shift, cM, cPX, cPY, cQX, cQY, cSA, cSD, cI):
    """one mimic_ec_mult_air(m, point, shift) (signature.py:176-190) into rows base .. base + 255; returns the sum"""
    if not 0 < m < 1 << NBITS:
        raise ValueError("scalar out of range")
    ps, q = shift, point
    for t in range(BLOCK):
        r = base + t
        cols[cM][r] = m >> t
        cols[cPX][r], cols[cPY][r] = ps
        if t <= NBITS and cQX is not None:
            cols[cQX][r], cols[cQY][r] = q                    # row 251 still holds 2^251 Q (step 250's doubling)
        if t < NBITS:
            d = (ps[0] - q[0]) % P
            if d == 0:
                raise ValueError("x collision in mimic_ec_mult_air")
            cols[cI][r] = inv(d)
            if (m >> t) & 1:
                cols[cSA][r] = (ps[1] - q[1]) * cols[cI][r] % P
                ps = ec_add(ps, q)
            if cSD is not None:
                cols[cSD][r] = (3 * q[0] * q[0] + 1) * inv(2 * q[1]) % P
            q = ec_double(q)
    return ps


def gen_trace(log_n, sigs):
    """sigs: N/256 tuples (msg_hash, r, w, (key x, key y)) with verify() == True.  -> columns [25][N].
    Raises ValueError where the reference's verify asserts or returns False."""
    n = 1 << log_n
    nb = n // BLOCK
    assert len(sigs) == nb
    cols = [[0] * n for _ in range(N_COLS)]
    sx, sy = SHIFT_POINT
    sums = []
    for b, (z, r, w, key) in enumerate(sigs):
        base = BLOCK * b
        if (key[1] * key[1] - key[0] ** 3 - key[0] - BETA) % P:
            raise ValueError("public key is not on the curve")
        zg = _walk(cols, base, z, EC_GEN, MINUS_SHIFT_POINT, AM, APX, APY, None, None, ASA, None, AI)
        rq = _walk(cols, base, r, key, SHIFT_POINT, BM, BPX, BPY, BQX, BQY, BSA, BSD, BI)
        d = (zg[0] - rq[0]) % P
        if d == 0:
            raise ValueError("x collision in ec_add(zG, rQ)")
        last = base + BLOCK - 1
        cols[AI][last] = inv(d)
        cols[ASA][last] = (zg[1] - rq[1]) * cols[AI][last] % P
        sums.append(ec_add(zg, rq))
        for t in range(BLOCK):
            cols[T1][base + t] = r
        cols[V1][base] = inv(z * r)
    for b in range(nb):
        base = BLOCK * b
        z, r, w, key = sigs[(b - 1) % nb]
        wb = _walk(cols, base, w, sums[(b - 1) % nb], SHIFT_POINT, CM, CPX, CPY, CQX, CQY, CSA, CSD, CI)
        d = (wb[0] - sx) % P
        if d == 0:
            raise ValueError("x collision in ec_add(wB, -shift)")
        last = base + BLOCK - 1
        cols[CI][last] = inv(d)
        cols[CSA][last] = (wb[1] + sy) * cols[CI][last] % P
        if (cols[CSA][last] ** 2 - wb[0] - sx - r) % P:
            raise ValueError("signature %d does not verify" % ((b - 1) % nb))
        for t in range(BLOCK):
            cols[T2][base + t] = r
        cols[V2][base] = inv(w)
    return cols


class EcdsaAir:
    kind = KIND
    n_alpha = N_ALPHA

    def __init__(self, log_n, pub):
        """pub = [(msg_hash, key x)] of the N/256 signatures"""
        if log_n < 9:
            raise ValueError("log_n too small")
        self.log_n, self.n = log_n, 1 << log_n
        self.pub = [(m % P, k % P) for m, k in pub]
        if len(self.pub) != self.n // BLOCK:
            raise ValueError("public input must list one (msg_hash, key) pair per block")
        if not all(0 < m < 1 << NBITS for m, _k in self.pub):
            raise ValueError("public scalar out of range")              # signature.py:219-227
        # the public columns: F(w_nb^b) = value of block b, as polynomials in x of degree < nb
        self.fm_coef = ontt.ntt([m for m, _k in self.pub], inverse=True)
        self.fk_coef = ontt.ntt([k for _m, k in self.pub], inverse=True)
        self.w256 = root_of_unity(8)
        gx, gy = doubled_generator()
        self.gx, self.gy = gx, gy
        self.gx_coef, self.gy_coef = ontt.ntt(gx, inverse=True), ontt.ntt(gy, inverse=True)

    # ---- protocol hooks (same set as stark.Air)
    def seed(self, n_queries):
        return (b"ecdsa-builtin" + self.log_n.to_bytes(4, "little") + n_queries.to_bytes(4, "little")
                + b"".join(ser(m) + ser(k) for m, k in self.pub))

    def header(self, n_queries, n_folds):
        return b"".join([stark.MAGIC, self.kind.to_bytes(4, "little"), self.log_n.to_bytes(4, "little"), (0).to_bytes(4, "little"),
                         n_queries.to_bytes(4, "little"), n_folds.to_bytes(4, "little")] + [ser(m) + ser(k) for m, k in self.pub])

    def statement(self):
        return {"log_n": self.log_n, "air": "ecdsa", "msgs": [m for m, _k in self.pub], "keys": [k for _m, k in self.pub]}

    def periodic_lde(self):
        g256 = pow(stark.GEN, self.n // BLOCK, P)
        gx_lde = ontt.lde(self.gx, stark.LOG_BLOWUP, g256)              # [j][i mod 256]
        gy_lde = ontt.lde(self.gy, stark.LOG_BLOWUP, g256)
        pad = [0] * (self.n - len(self.pub))
        fm_lde = ontt.lde(ontt.ntt(self.fm_coef + pad), stark.LOG_BLOWUP, stark.GEN)     # [j][i]: F at the LDE points
        fk_lde = ontt.lde(ontt.ntt(self.fk_coef + pad), stark.LOG_BLOWUP, stark.GEN)
        return lambda j, i: (gx_lde[j][i % BLOCK], gy_lde[j][i % BLOCK], fm_lde[j][i], fk_lde[j][i])

    def periodic_at(self, x):
        u = pow(x, self.n // BLOCK, P)
        hx = hy = 0
        for cx, cy in zip(reversed(self.gx_coef), reversed(self.gy_coef)):
            hx, hy = (hx * u + cx) % P, (hy * u + cy) % P
        fm = fk = 0
        for cm, ck in zip(reversed(self.fm_coef), reversed(self.fk_coef)):
            fm, fk = (fm * x + cm) % P, (fk * x + ck) % P
        return hx, hy, fm, fk

    def inv_zerofiers(self, x):
        u = pow(x, self.n // BLOCK, P)
        z_all = (pow(x, self.n, P) - 1) % P
        tail4 = 1
        for t in range(NBITS, BLOCK - 1):
            tail4 = tail4 * (u - pow(self.w256, t, P)) % P
        e_last = (u - pow(self.w256, BLOCK - 1, P)) % P
        tail5 = tail4 * e_last % P
        iz_all = inv(z_all)
        return {
            "step": tail5 * iz_all % P,            # rows t <= 250 of every block
            "hold": inv(tail4),                    # rows 251 .. 254
            "zero": inv(tail5),                    # rows 251 .. 255
            "first": inv(u - 1),                   # t = 0
            "last": inv(e_last),                   # t = 255
            "thold": e_last * iz_all % P,          # every row but t = 255
        }

    @staticmethod
    def _lane(a, M, Mn, PX, PY, PXn, PYn, QX, QY, SA, I):
        """the five step constraints every lane has (bit, slope, x, y, distinctness), already weighted"""
        bit = (M - 2 * Mn) % P
        nb = (1 - bit) % P
        c1 = bit * (bit - 1)
        c2 = bit * (SA * (PX - QX) - (PY - QY))
        c3 = bit * (SA * SA - PX - QX - PXn) + nb * (PXn - PX)
        c4 = bit * (SA * (PX - PXn) - PY - PYn) + nb * (PYn - PY)
        c5 = I * (PX - QX) - 1
        return (a[0] * c1 + a[1] * c2 + a[2] * c3 + a[3] * c4 + a[4] * c5) % P

    @staticmethod
    def _double(a, QX, QY, QXn, QYn, SD):
        d1 = 2 * SD * QY - 3 * QX * QX - 1
        d2 = SD * SD - 2 * QX - QXn
        d3 = SD * (QX - QXn) - QY - QYn
        return (a[0] * d1 + a[1] * d2 + a[2] * d3) % P

    def composition_per(self, cur, nxt, per, iz, a):
        gx, gy, fm, fk = per
        sx, sy = SHIFT_POINT
        c, n = cur, nxt
        # ---- lane A: z G, shift -S, point = periodic (gx, gy)                                  alpha 0 .. 9
        step = self._lane(a[0:5], c[AM], n[AM], c[APX], c[APY], n[APX], n[APY], gx, gy, c[ASA], c[AI])
        hold = a[5] * (n[APX] - c[APX]) + a[6] * (n[APY] - c[APY])
        zero = a[7] * c[AM]
        first = a[8] * (c[APX] - MINUS_SHIFT_POINT[0]) + a[9] * (c[APY] - MINUS_SHIFT_POINT[1])
        # ---- lane B: r Q, shift S                                                              alpha 10 .. 23
        step += self._lane(a[10:15], c[BM], n[BM], c[BPX], c[BPY], n[BPX], n[BPY], c[BQX], c[BQY], c[BSA], c[BI])
        step += self._double(a[15:18], c[BQX], c[BQY], n[BQX], n[BQY], c[BSD])
        hold += a[18] * (n[BPX] - c[BPX]) + a[19] * (n[BPY] - c[BPY])
        zero += a[20] * c[BM]
        first += a[21] * (c[BPX] - sx) + a[22] * (c[BPY] - sy)
        first += a[23] * (c[BQY] * c[BQY] - c[BQX] * c[BQX] * c[BQX] - c[BQX] - BETA)              # the key is on the curve
        # ---- lane C: w (zG + rQ), shift S                                                      alpha 24 .. 36
        step += self._lane(a[24:29], c[CM], n[CM], c[CPX], c[CPY], n[CPX], n[CPY], c[CQX], c[CQY], c[CSA], c[CI])
        step += self._double(a[29:32], c[CQX], c[CQY], n[CQX], n[CQY], c[CSD])
        hold += a[32] * (n[CPX] - c[CPX]) + a[33] * (n[CPY] - c[CPY])
        zero += a[34] * c[CM]
        first += a[35] * (c[CPX] - sx) + a[36] * (c[CPY] - sy)
        # ---- row 255: ec_add(zG, rQ) -> lane C's point on the next row; ec_add(wB, -S).x == r    alpha 37 .. 44
        dab = c[APX] - c[BPX]
        last = a[37] * (c[AI] * dab - 1) + a[38] * (c[ASA] * dab - (c[APY] - c[BPY]))
        last += a[39] * (n[CQX] - (c[ASA] * c[ASA] - c[APX] - c[BPX]))
        last += a[40] * (n[CQY] - (c[ASA] * (c[APX] - n[CQX]) - c[APY]))
        dcs = c[CPX] - sx
        last += a[41] * (c[CI] * dcs - 1) + a[42] * (c[CSA] * dcs - (c[CPY] + sy))
        last += a[43] * (c[CSA] * c[CSA] - c[CPX] - sx - c[T2])
        last += a[44] * (n[T2] - c[T1])
        # ---- row 0: r into its carrier, scalars non-zero                                       alpha 45 .. 47
        first += a[45] * (c[T1] - c[BM]) + a[46] * (c[V1] * c[AM] % P * c[BM] - 1) + a[47] * (c[V2] * c[CM] - 1)
        # ---- carriers hold inside a block                                                      alpha 48, 49
        thold = a[48] * (n[T1] - c[T1]) + a[49] * (n[T2] - c[T2])
        # ---- public input on the block-start rows: the message and the key's x                 alpha 50, 51
        first += a[50] * (c[AM] - fm) + a[51] * (c[BQX] - fk)
        acc = (step % P * iz["step"] + hold % P * iz["hold"] + zero % P * iz["zero"] + first % P * iz["first"]
               + last % P * iz["last"] + thold % P * iz["thold"])
        return acc % P


def public_of(sigs):
    return [(z, key[0]) for z, _r, _w, key in sigs]


def prove(log_n, sigs, n_queries=30, corrupt=None, debug=None):
    """proof bytes for N/256 valid signatures.  corrupt = (col, row, delta): tamper with one cell (negative tests)"""
    cols = gen_trace(log_n, sigs)
    if corrupt:
        c, r, d = corrupt
        cols[c][r] = (cols[c][r] + d) % P
    return stark.prove_air(EcdsaAir(log_n, public_of(sigs)), cols, n_queries, debug)


def make_signatures(count, seed=1):
    """count valid (msg_hash, r, w, key point) tuples made with the oracle's signer (oracle/ecdsa.py, RFC 6979)"""
    import random
    from . import ecdsa as oe
    rng = random.Random(seed)
    out = []
    while len(out) < count:
        priv = rng.randrange(1, 1 << 250)
        msg = rng.randrange(1, 1 << NBITS)
        r, s = oe.sign(msg, priv)
        key = oe.private_key_to_ec_point_on_stark_curve(priv)
        assert oe.verify(msg, r, s, key)
        out.append((msg, r, oe.inv_mod_curve_size(s), key))
    return out
