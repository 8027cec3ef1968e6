"""STARK prover + verifier, plain Python ints: the protocol (AIR-generic: prove_air / verify) and the Pedersen hash-chain
AIR; the second AIR (ECDSA builtin) lives in stark_ecdsa.py.  TEST INFRASTRUCTURE.

PARITY UNPINNED: the reference repository contains no prover, verifier or AIR; the
protocol below is this repo's own (DESIGN.md "Protocol").  What IS pinned by the reference is the
witness semantics: every trace row is one step of signature.py:300-318 (pedersen_hash_as_point), with
its `point.x != pt.x` assertion turned into the inverse column.

Trace: N = 2^log_n rows, 5 lanes x 5 columns (X, Y, S, M, I); one Pedersen hash instance = 512 rows
(2 elements x 256 rows: 252 bit steps + 4 padding rows).  Instances of a lane are chained in segments of
2^chain_log hashes: the first element of an instance is the previous instance's result, except at a
segment start where it is the public seed x0[lane].  Public output: X at the last row, per lane.
"""
import hashlib

from .params import CONSTANT_POINTS, FIELD_PRIME as P, SHIFT_POINT, R_MOD_P, root_of_unity
from . import ntt as ontt

LANES = 5
COLS_PER_LANE = 5
N_COLS = LANES * COLS_PER_LANE
N_CONSTRAINTS = 13
LOG_BLOWUP = 3
BLOWUP = 8
GEN = 3
RINV = pow(R_MOD_P, -1, P)
MAGIC = b"SPGP"
VERSION = 1
LAST_LAYER_MAX = 64
CANON_BITS = 251          # scalars are unpacked into 251 bits (M = 0 from row 251 on): unique decomposition mod p
MIN_QUERIES = 30          # the protocol's query count; verify() rejects proofs carrying fewer unless told otherwise


def inv(a):
    return pow(a % P, -1, P)


# ------------------------------------------------------------------ serialisation / hashing / channel
def ser(v):
    """field element -> 32 bytes: Montgomery representation (v * 2^256 mod p), big-endian."""
    return (v * R_MOD_P % P).to_bytes(32, "big")


def deser(b):
    m = int.from_bytes(b, "big")
    if m >= P:
        raise ValueError("non-canonical field element")
    return m * RINV % P


def H(data):
    return hashlib.blake2s(data).digest()


class Channel:
    def __init__(self, seed):
        self.state = H(b"spg-stark-v1" + seed)
        self.counter = 0

    def absorb(self, data):
        self.state = H(self.state + data)
        self.counter = 0

    def draw(self):
        out = H(self.state + self.counter.to_bytes(8, "little"))
        self.counter += 1
        return out

    def draw_felt(self):
        return int.from_bytes(self.draw(), "little") & ((1 << 251) - 1)

    def draw_index(self, n):
        return int.from_bytes(self.draw()[:8], "little") % n


def merkle_levels(leaves):
    """leaves: list of 32-byte hashes (power of two).  Returns [level0 = leaves, level1, ..., [root]]."""
    levels = [leaves]
    while len(levels[-1]) > 1:
        prev = levels[-1]
        levels.append([H(prev[2 * i] + prev[2 * i + 1]) for i in range(len(prev) // 2)])
    return levels


def merkle_path(levels, idx):
    path = []
    for lvl in levels[:-1]:
        path.append(lvl[idx ^ 1])
        idx >>= 1
    return path


def merkle_root_from_path(leaf_hash, idx, path):
    h = leaf_hash
    for sib in path:
        h = H(sib + h) if idx & 1 else H(h + sib)
        idx >>= 1
    return h


def table_leaves(table, rows):
    """table[j][c][i] (8 cosets, ncols columns, `rows` rows) -> list of `rows` leaf byte strings; leaf
    j*(rows/8) + i' holds rows i' + k*rows/8, k = 0..7, all columns each."""
    g = rows // 8
    out = []
    for j in range(BLOWUP):
        for ip in range(g):
            out.append(b"".join(ser(col[ip + k * g]) for k in range(8) for col in table[j]))
    return out


# ------------------------------------------------------------------ AIR
def periodic_points():
    """PX, PY over one 512-row instance (0 on padding rows)."""
    px, py = [0] * 512, [0] * 512
    for e in range(2):
        for t in range(252):
            px[256 * e + t], py[256 * e + t] = CONSTANT_POINTS[2 + 252 * e + t]
    return px, py


def gen_trace(log_n, chain_log, x0, ys, unpack_override=None):
    """x0: LANES seeds; ys[lane][instance] second hash inputs.  Returns (columns [25][N], outputs [LANES]).
    unpack_override (negative tests only): {(lane, instance, element): integer w} walks the bits of w (w = v mod p,
    e.g. v + p) instead of the canonical v -- the cheating witness the canonical-unpacking constraint must reject."""
    n = 1 << log_n
    inst = n // 512
    cols = [[0] * n for _ in range(N_COLS)]
    outs = []
    for l in range(LANES):
        X, Y, S, M, I = (cols[5 * l + k] for k in range(5))
        prev = None
        for q in range(inst):
            a = x0[l] if q % (1 << chain_log) == 0 else prev
            elems = (a, ys[l][q])
            pt_sum = SHIFT_POINT
            for e in range(2):
                v = elems[e]
                assert 0 <= v < P
                if unpack_override and (l, q, e) in unpack_override:
                    assert unpack_override[(l, q, e)] % P == v
                    v = unpack_override[(l, q, e)]
                else:
                    assert v < (1 << CANON_BITS), "hash input >= 2^251: outside the AIR's canonical range"
                for t in range(256):
                    r = 512 * q + 256 * e + t
                    X[r], Y[r], M[r] = pt_sum[0], pt_sum[1], (v >> t) % P
                    if t < 252:
                        cx, cy = CONSTANT_POINTS[2 + 252 * e + t]
                        d = (pt_sum[0] - cx) % P
                        assert d != 0, "Unhashable input."
                        I[r] = inv(d)
                        if (v >> t) & 1:
                            s = (pt_sum[1] - cy) * I[r] % P
                            S[r] = s
                            nx = (s * s - pt_sum[0] - cx) % P
                            pt_sum = (nx, (s * (pt_sum[0] - nx) - pt_sum[1]) % P)
            prev = pt_sum[0]
        outs.append(prev)
    return cols, outs


class Air:
    """Constraint evaluation at an arbitrary point x (given current / next row values and the periodic
    values at x) -- used on the LDE domain by the prover and at the OODS point by the verifier."""

    kind = 1                                   # proof header VERSION field
    n_alpha = LANES * N_CONSTRAINTS

    def __init__(self, log_n, chain_log, x0, outs):
        self.log_n, self.n = log_n, 1 << log_n
        self.chain_log = chain_log
        self.seg = 512 << chain_log
        assert self.seg <= self.n
        self.x0, self.outs = x0, outs
        self.w256 = root_of_unity(8)
        self.w512 = root_of_unity(9)
        self.wseg = root_of_unity(9 + chain_log)
        self.wn = root_of_unity(log_n)
        px, py = periodic_points()
        self.px_coef = ontt.ntt(px, inverse=True)
        self.py_coef = ontt.ntt(py, inverse=True)

    def inv_zerofiers(self, x):
        n = self.n
        u256, u512, useg = pow(x, n // 256, P), pow(x, n // 512, P), pow(x, n // self.seg, P)
        z_all = (pow(x, n, P) - 1) % P
        e_step = (u256 - pow(self.w256, 255, P)) % P
        z_pad = 1
        for k in range(252, 256):
            z_pad = z_pad * (u256 - pow(self.w256, k, P)) % P
        # c6 (M = 0) also covers row 251: the scalar is unpacked into 251 bits only, so the integer the bits spell
        # is below 2^251 < p and therefore THE canonical representative of M_0 (signature.py:307 `0 <= x < p`
        # uses the integer; a 252-bit unpacking would also accept x + p for x < 2^252 - p)
        z_zero = z_pad * (u256 - pow(self.w256, CANON_BITS, P)) % P
        iz_all = inv(z_all)
        return {
            "step": e_step * iz_all % P,
            "act": z_pad * iz_all % P,
            "pad": inv(z_zero),
            "mid": inv(u512 - pow(self.w512, 255, P)),
            "link": (useg - inv(self.wseg)) * inv(u512 - pow(self.w512, 511, P)) % P,
            "inst0": inv(u512 - 1),
            "seg0": inv(useg - 1),
            "last": inv(x - pow(self.wn, self.n - 1, P)),
        }

    def periodic_at(self, x):
        u = pow(x, self.n // 512, P)
        hx = hy = 0
        for cx, cy in zip(reversed(self.px_coef), reversed(self.py_coef)):
            hx, hy = (hx * u + cx) % P, (hy * u + cy) % P
        return hx, hy

    # ---- hooks of the AIR-generic prover / verifier below (the ECDSA-builtin AIR of stark_ecdsa.py has the same set)
    def seed(self, n_queries):
        return public_seed(self.log_n, self.chain_log, n_queries, self.x0, self.outs)

    def header(self, n_queries, n_folds):
        return b"".join([MAGIC, self.kind.to_bytes(4, "little"), self.log_n.to_bytes(4, "little"),
                         self.chain_log.to_bytes(4, "little"), n_queries.to_bytes(4, "little"), n_folds.to_bytes(4, "little")]
                        + [ser(v) for v in self.x0] + [ser(v) for v in self.outs])

    def periodic_lde(self):
        """-> per(j, i): the periodic columns' values at LDE point (coset j, row i)"""
        px512, py512 = periodic_points()
        g512 = pow(GEN, self.n // 512, P)
        px_lde = ontt.lde(px512, LOG_BLOWUP, g512)                        # [j][i mod 512]
        py_lde = ontt.lde(py512, LOG_BLOWUP, g512)
        return lambda j, i: (px_lde[j][i % 512], py_lde[j][i % 512])

    def composition_per(self, cur, nxt, per, iz, alpha_pows):
        return self.composition(cur, nxt, per[0], per[1], iz, alpha_pows)

    def statement(self):
        return {"log_n": self.log_n, "chain_log": self.chain_log, "x0": self.x0, "outs": self.outs}

    def composition(self, cur, nxt, px, py, iz, alpha_pows):
        """cur / nxt: 25 values at x and x * w_N; returns CP(x)."""
        sx, sy = SHIFT_POINT
        acc = 0
        for l in range(LANES):
            X, Y, S, M, I = cur[5 * l:5 * l + 5]
            Xn, Yn, _Sn, Mn, _In = nxt[5 * l:5 * l + 5]
            a = alpha_pows[13 * l:13 * l + 13]
            bit = (M - 2 * Mn) % P
            nb = (1 - bit) % P
            c1 = bit * (bit - 1)
            c2 = bit * (S * (X - px) - (Y - py))
            c3 = bit * (S * S - X - px - Xn) + nb * (Xn - X)
            c4 = bit * (S * (X - Xn) - Y - Yn) + nb * (Yn - Y)
            c5 = I * (X - px) - 1
            c6 = M
            c7, c8 = Xn - X, Yn - Y
            c9 = Mn - X
            c10, c11 = X - sx, Y - sy
            c12 = M - self.x0[l]
            c13 = X - self.outs[l]
            acc += (a[0] * c1 + a[1] * c2 + a[2] * c3 + a[3] * c4) % P * iz["step"]
            acc += a[4] * c5 % P * iz["act"] + a[5] * c6 % P * iz["pad"]
            acc += (a[6] * c7 + a[7] * c8) % P * iz["mid"] + a[8] * c9 % P * iz["link"]
            acc += (a[9] * c10 + a[10] * c11) % P * iz["inst0"] + a[11] * c12 % P * iz["seg0"]
            acc += a[12] * c13 % P * iz["last"]
        return acc % P


def lde_point(log_n, j, i):
    return GEN * pow(root_of_unity(log_n + LOG_BLOWUP), j + 8 * i, P) % P


def fri_layer_sizes(log_n):
    """rows-per-coset N_l of every layer, layer 0 = N, folding by 8 while N_l > 64."""
    sizes = [1 << log_n]
    while sizes[-1] > LAST_LAYER_MAX:
        sizes.append(sizes[-1] // 8)
    return sizes


def fold8(vals, x, beta):
    """vals[k] = P(x * w_8^k), k < 8  ->  sum_m beta^m P_m(x^8)  where P(t) = sum_m t^m P_m(t^8)."""
    zeta_inv = inv(root_of_unity(3))
    t = beta * inv(x) % P
    acc, tm = 0, 1
    for m in range(8):
        s = 0
        for k in range(8):
            s += pow(zeta_inv, m * k, P) * vals[k]
        acc += tm * (s % P)
        tm = tm * t % P
    return acc * inv(8) % P


def public_seed(log_n, chain_log, n_queries, x0, outs):
    return (log_n.to_bytes(4, "little") + chain_log.to_bytes(4, "little") + n_queries.to_bytes(4, "little")
            + b"".join(ser(v) for v in x0) + b"".join(ser(v) for v in outs))


def deep_quotient(tvals, hvals, x, z, zw, z4, oods, gamma_pows):
    """tvals: 25 trace values at x, hvals: 4 composition-chunk values at x."""
    tz, tzw, hz = oods[:25], oods[25:50], oods[50:54]
    a = sum(gamma_pows[c] * (tvals[c] - tz[c]) for c in range(25)) % P
    b = sum(gamma_pows[25 + c] * (tvals[c] - tzw[c]) for c in range(25)) % P
    c_ = sum(gamma_pows[50 + m] * (hvals[m] - hz[m]) for m in range(4)) % P
    return (a * inv(x - z) + b * inv(x - zw) + c_ * inv(x - z4)) % P


# ------------------------------------------------------------------ prover
def prove(log_n, chain_log, x0, ys, n_queries=30, corrupt=None):
    """Returns proof bytes.  `corrupt` = (col, row, delta) tampers with one trace cell after witness
    generation (negative tests)."""
    n = 1 << log_n
    cols, outs = gen_trace(log_n, chain_log, x0, ys)
    if corrupt:
        c, r, d = corrupt
        cols[c][r] = (cols[c][r] + d) % P
    return prove_trace(log_n, chain_log, x0, outs, cols, n_queries)


def prove_trace(log_n, chain_log, x0, outs, cols, n_queries=30, debug=None):
    return prove_air(Air(log_n, chain_log, x0, outs), cols, n_queries, debug)


def prove_air(air, cols, n_queries=30, debug=None):
    """The protocol for any AIR object with the hook set of class Air (25 columns, mask {x, x w_N}, composition degree
    < 4N): the Pedersen hash chain above, the ECDSA builtin of stark_ecdsa.py."""
    log_n = air.log_n
    n = 1 << log_n
    ch = Channel(air.seed(n_queries))
    wn = root_of_unity(log_n)
    # 1. trace LDE + commitment
    lde_cols = [ontt.lde(c, LOG_BLOWUP, GEN) for c in cols]               # [c][j][i]
    t_table = [[lde_cols[c][j] for c in range(N_COLS)] for j in range(BLOWUP)]
    t_leaves = table_leaves(t_table, n)
    t_levels = merkle_levels([H(x) for x in t_leaves])
    ch.absorb(t_levels[-1][0])
    # 2. composition on the cosets j = 0, 2, 4, 6  (= the coset g <w_4N>)
    alpha = ch.draw_felt()
    apows = [pow(alpha, k, P) for k in range(air.n_alpha)]
    per = air.periodic_lde()
    cp = [0] * (4 * n)                                                    # index e' = j/2 + 4 i
    for j in range(0, 8, 2):
        for i in range(n):
            x = lde_point(log_n, j, i)
            cur = [t_table[j][c][i] for c in range(N_COLS)]
            nxt = [t_table[j][c][(i + 1) % n] for c in range(N_COLS)]
            iz = air.inv_zerofiers(x)
            cp[j // 2 + 4 * i] = air.composition_per(cur, nxt, per(j, i), iz, apows)
    # interpolate CP on g <w_4N>, split into 4 chunks of degree < N
    coef = ontt.ntt(cp, inverse=True)
    ginv = inv(GEN)
    coef = [c * pow(ginv, k, P) % P for k, c in enumerate(coef)]
    h_coef = [coef[m::4] for m in range(4)]
    if debug is not None:
        debug["cp"] = cp
        debug["h_coef"] = h_coef
    # LDE of the chunks: evaluate H_m at the LDE points themselves
    h_lde = []
    for m in range(4):
        evals = ontt.ntt(h_coef[m])                                       # values on <w_N>
        h_lde.append(ontt.lde(evals, LOG_BLOWUP, GEN))
    h_table = [[h_lde[m][j] for m in range(4)] for j in range(BLOWUP)]
    h_leaves = table_leaves(h_table, n)
    h_levels = merkle_levels([H(x) for x in h_leaves])
    ch.absorb(h_levels[-1][0])
    # 3. out-of-domain sampling
    z = ch.draw_felt()
    zw, z4 = z * wn % P, pow(z, 4, P)

    def horner(cf, pt):
        acc = 0
        for c in reversed(cf):
            acc = (acc * pt + c) % P
        return acc
    t_coef = [ontt.ntt(c, inverse=True) for c in cols]
    oods = [horner(c, z) for c in t_coef] + [horner(c, zw) for c in t_coef] + [horner(c, z4) for c in h_coef]
    ch.absorb(b"".join(ser(v) for v in oods))
    # 4. DEEP quotient on the whole LDE domain
    gamma = ch.draw_felt()
    gpows = [pow(gamma, k, P) for k in range(54)]
    layer = [[deep_quotient([t_table[j][c][i] for c in range(N_COLS)], [h_table[j][m][i] for m in range(4)],
                            lde_point(log_n, j, i), z, zw, z4, oods, gpows) for i in range(n)] for j in range(BLOWUP)]
    if debug is not None:
        debug["oods"] = oods
        debug["deep"] = layer
    # 5. FRI
    sizes = fri_layer_sizes(log_n)
    fri_tables, fri_levels, betas = [], [], []
    g_l, cur_log = GEN, log_n
    for li in range(1, len(sizes)):
        beta = ch.draw_felt()
        betas.append(beta)
        rows, grp = sizes[li - 1], sizes[li]
        w = root_of_unity(cur_log + LOG_BLOWUP)
        new = [[fold8([layer[j][ip + k * grp] for k in range(8)], g_l * pow(w, j + 8 * ip, P) % P, beta)
                for ip in range(grp)] for j in range(BLOWUP)]
        layer = new
        g_l, cur_log = pow(g_l, 8, P), cur_log - 3
        tbl = [[layer[j]] for j in range(BLOWUP)]
        lv = merkle_levels([H(x) for x in table_leaves(tbl, grp)])
        fri_tables.append(tbl)
        fri_levels.append(lv)
        ch.absorb(lv[-1][0])
    # last layer: interpolate on g_l <w_{8 N_last}>, natural index e = j + 8 i
    n_last = sizes[-1]
    flat = [0] * (8 * n_last)
    for j in range(BLOWUP):
        for i in range(n_last):
            flat[j + 8 * i] = layer[j][i]
    lc = ontt.ntt(flat, inverse=True)
    gli = inv(g_l)
    lc = [c * pow(gli, k, P) % P for k, c in enumerate(lc)]
    if any(lc[n_last:]):
        raise ValueError("trace does not satisfy the AIR: FRI last layer has degree >= %d" % n_last)
    last_coef = lc[:n_last]
    ch.absorb(b"".join(ser(v) for v in last_coef))
    if debug is not None:
        debug["betas"] = betas
        debug["last_coef"] = last_coef
    # 6. queries
    out = [air.header(n_queries, len(sizes) - 1)]
    out += [t_levels[-1][0], h_levels[-1][0]] + [ser(v) for v in oods]
    out += [lv[-1][0] for lv in fri_levels] + [ser(v) for v in last_coef]
    for _ in range(n_queries):
        idx = ch.draw_index(n)
        out += [t_leaves[idx]] + merkle_path(t_levels, idx) + [h_leaves[idx]] + merkle_path(h_levels, idx)
        j, ip = idx // (n // 8), idx % (n // 8)
        for li in range(1, len(sizes)):
            rows = sizes[li]
            g8 = rows // 8
            ip = ip % g8
            out += [table_leaves_one(fri_tables[li - 1], rows, j, ip)] + merkle_path(fri_levels[li - 1], j * g8 + ip)
    return b"".join(out)


def table_leaves_one(table, rows, j, ip):
    g = rows // 8
    return b"".join(ser(col[ip + k * g]) for k in range(8) for col in table[j])


# ------------------------------------------------------------------ verifier
class ProofError(Exception):
    pass


class _Reader:
    def __init__(self, b):
        self.b, self.o = b, 0

    def take(self, n):
        if self.o + n > len(self.b):
            raise ProofError("proof truncated")
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def u32(self):
        return int.from_bytes(self.take(4), "little")

    def felt(self):
        try:
            return deser(self.take(32))
        except ValueError as e:
            raise ProofError(str(e))


def verify(proof, min_queries=MIN_QUERIES):
    """Returns the public statement {log_n, chain_log, x0, outs} if the proof is valid, raises ProofError otherwise.
    The query count is read from the proof header; proofs with fewer than `min_queries` queries are rejected
    (blowup 8, no grinding: each query is worth 3 bits)."""
    rd = _Reader(proof)
    if rd.take(4) != MAGIC:
        raise ProofError("bad header")
    kind = rd.u32()
    if kind not in (VERSION, 2):
        raise ProofError("bad header")
    log_n, chain_log, n_queries, n_folds = rd.u32(), rd.u32(), rd.u32(), rd.u32()
    if not (9 <= log_n <= 23) or n_queries < 1:
        raise ProofError("bad parameters")
    if kind == VERSION and 512 << chain_log > 1 << log_n:
        raise ProofError("bad parameters")
    if kind == 2 and chain_log != 0:
        raise ProofError("bad parameters")
    if n_queries < min_queries:
        raise ProofError("proof carries %d queries, fewer than the %d required" % (n_queries, min_queries))
    sizes = fri_layer_sizes(log_n)
    if n_folds != len(sizes) - 1:
        raise ProofError("bad layer count")
    n = 1 << log_n
    if kind == VERSION:
        x0 = [rd.felt() for _ in range(LANES)]
        outs = [rd.felt() for _ in range(LANES)]
        air = Air(log_n, chain_log, x0, outs)
    else:                                          # the ECDSA-builtin AIR: public (msg_hash, key x) of every block
        from .stark_ecdsa import EcdsaAir
        try:
            air = EcdsaAir(log_n, [(rd.felt(), rd.felt()) for _ in range(n >> 8)])
        except ValueError as e:
            raise ProofError(str(e))
    root_t, root_h = rd.take(32), rd.take(32)
    oods = [rd.felt() for _ in range(54)]
    fri_roots = [rd.take(32) for _ in range(n_folds)]
    last_coef = [rd.felt() for _ in range(sizes[-1])]
    wn = root_of_unity(log_n)
    ch = Channel(air.seed(n_queries))
    ch.absorb(root_t)
    alpha = ch.draw_felt()
    apows = [pow(alpha, k, P) for k in range(air.n_alpha)]
    ch.absorb(root_h)
    z = ch.draw_felt()
    zw, z4 = z * wn % P, pow(z, 4, P)
    ch.absorb(b"".join(ser(v) for v in oods))
    # composition consistency at z
    cpz = air.composition_per(oods[:25], oods[25:50], air.periodic_at(z), air.inv_zerofiers(z), apows)
    if cpz != sum(pow(z, m, P) * oods[50 + m] for m in range(4)) % P:
        raise ProofError("composition polynomial mismatch at the out-of-domain point")
    gamma = ch.draw_felt()
    gpows = [pow(gamma, k, P) for k in range(54)]
    betas = []
    for li in range(n_folds):
        betas.append(ch.draw_felt())
        ch.absorb(fri_roots[li])
    ch.absorb(b"".join(ser(v) for v in last_coef))
    for _ in range(n_queries):
        idx = ch.draw_index(n)
        t_leaf = rd.take(8 * N_COLS * 32)
        t_path = [rd.take(32) for _ in range(log_n)]
        h_leaf = rd.take(8 * 4 * 32)
        h_path = [rd.take(32) for _ in range(log_n)]
        if merkle_root_from_path(H(t_leaf), idx, t_path) != root_t:
            raise ProofError("trace decommitment failed")
        if merkle_root_from_path(H(h_leaf), idx, h_path) != root_h:
            raise ProofError("composition decommitment failed")
        j, ip = idx // (n // 8), idx % (n // 8)
        try:
            tv = [[deser(t_leaf[32 * (k * N_COLS + c):32 * (k * N_COLS + c) + 32]) for c in range(N_COLS)] for k in range(8)]
            hv = [[deser(h_leaf[32 * (k * 4 + m):32 * (k * 4 + m) + 32]) for m in range(4)] for k in range(8)]
        except ValueError as e:
            raise ProofError(str(e))
        grp = n // 8
        vals = [deep_quotient(tv[k], hv[k], lde_point(log_n, j, ip + k * grp), z, zw, z4, oods, gpows) for k in range(8)]
        g_l, cur_log = GEN, log_n
        for li in range(1, len(sizes)):
            x = g_l * pow(root_of_unity(cur_log + LOG_BLOWUP), j + 8 * ip, P) % P
            v = fold8(vals, x, betas[li - 1])
            g_l, cur_log = pow(g_l, 8, P), cur_log - 3
            rows = sizes[li]
            g8 = rows // 8
            slot, ip2 = ip // g8, ip % g8
            leaf = rd.take(8 * 32)
            path = [rd.take(32) for _ in range(cur_log)]
            if merkle_root_from_path(H(leaf), j * g8 + ip2, path) != fri_roots[li - 1]:
                raise ProofError("FRI layer %d decommitment failed" % li)
            try:
                vals = [deser(leaf[32 * k:32 * k + 32]) for k in range(8)]
            except ValueError as e:
                raise ProofError(str(e))
            if vals[slot] != v:
                raise ProofError("FRI layer %d folding mismatch" % li)
            ip = ip2
        # `vals` is the opened group of the last layer; check every element against the polynomial
        rows = sizes[-1]
        g8 = rows // 8
        w = root_of_unity(cur_log + LOG_BLOWUP)
        for k in range(8):
            x = g_l * pow(w, j + 8 * (ip + k * g8), P) % P
            acc = 0
            for c in reversed(last_coef):
                acc = (acc * x + c) % P
            if acc != vals[k]:
                raise ProofError("FRI last layer mismatch")
    if rd.o != len(proof):
        raise ProofError("trailing bytes")
    st = air.statement()
    st["n_queries"] = n_queries
    return st
