"""Host-side driver of the STARK prover (the role Stone's `cpu_air_prover` CLI plays next to cairo-run's
artefacts, src/starkware/cairo/lang/cairo_cmake_rules.cmake:72-110 -- the reference itself has no prover).

Single GPU:   Prover(ctx).prove_host(trace, ...) -> libspg's spg_prove (all stages sequenced in C++).
Multi GPU:    Prover(ctx, rank, world) -- one process per GPU; this module sequences the stage-level C-ABI
              (spg_stage_*) around torch.distributed collectives.  Sharding (DESIGN.md "Multi-GPU"):
              * columns are sharded for interpolation, ONE all-gather exchanges the coefficient columns;
              * each GPU then owns 8/world consecutive cosets of the evaluation domain for everything else
                (LDE, Merkle sub-trees, AIR, DEEP, FRI folds never leave a coset);
              * 32-byte sub-tree roots, the four composition chunks (N felts each) and the query openings are
                the only other traffic.
              The proof is byte-identical to the single-GPU one.

The stage backend is an interface (`GpuBackend` here); tests/ plugs a CPU stand-in behind the same driver to
exercise the multi-process logic with gloo.
"""
import ctypes as C
import hashlib

import numpy as np

from ._lib import FIELD_PRIME as P, SPG_DEVICE_PTRS, SPG_NO_SYNC, ints_to_limbs, limbs_to_ints

R_MOD_P = (1 << 256) % P
LANES, N_COLS, N_CONSTR, BLOWUP, LOG_BLOWUP, GEN = 5, 25, 13, 8, 3, 3
LAST_LAYER_MAX = 64
SPG_MONT_OUT = 4


# ------------------------------------------------------------------ host-side field / hash / channel helpers
def ser(v):
    return (v * R_MOD_P % P).to_bytes(32, "big")


def H(b):
    return hashlib.blake2s(b).digest()


def root_of_unity(log_n):
    return pow(GEN, (P - 1) >> log_n, P)


class Channel:
    def __init__(self, seed):
        self.state, self.counter = H(b"spg-stark-v1" + seed), 0

    def absorb(self, data):
        self.state, self.counter = H(self.state + data), 0

    def draw(self):
        out = H(self.state + self.counter.to_bytes(8, "little"))
        self.counter += 1
        return out

    def draw_felt(self):
        return int.from_bytes(self.draw(), "little") & ((1 << 251) - 1)

    def draw_index(self, n):
        return int.from_bytes(self.draw()[:8], "little") % n


def fri_log_rows(log_n):
    lr = [log_n]
    while (1 << lr[-1]) > LAST_LAYER_MAX:
        lr.append(lr[-1] - 3)
    return lr


def top_levels(roots):
    """Merkle levels above the per-GPU sub-tree roots."""
    levels = [list(roots)]
    while len(levels[-1]) > 1:
        prev = levels[-1]
        levels.append([H(prev[2 * i] + prev[2 * i + 1]) for i in range(len(prev) // 2)])
    return levels


def top_path(levels, owner):
    path, idx = [], owner
    for lvl in levels[:-1]:
        path.append(lvl[idx ^ 1])
        idx >>= 1
    return path


def composition_at(log_n, chain_log, x0, outs, alpha_pows, z, tz, tzw, const_points, shift):
    """CP(z) from the trace values at z and z*w (prover self-check; same formulas as csrc/air.cu)."""
    n, seg = 1 << log_n, 512 << chain_log
    inv = lambda a: pow(a % P, -1, P)   # noqa: E731
    u256, u512, useg = pow(z, n // 256, P), pow(z, n // 512, P), pow(z, n // seg, P)
    w256, w512, wseg = root_of_unity(8), root_of_unity(9), root_of_unity(9 + chain_log)
    iz_all = inv(pow(z, n, P) - 1)
    z_pad = 1
    for k in range(252, 256):
        z_pad = z_pad * (u256 - pow(w256, k, P)) % P
    z_zero = z_pad * (u256 - pow(w256, 251, P)) % P      # c6 (M = 0) from row 251 on: canonical 251-bit unpacking
    iz = [(u256 - pow(w256, 255, P)) * iz_all % P, z_pad * iz_all % P, inv(z_zero), inv(u512 - pow(w512, 255, P)),
          (useg - inv(wseg)) * inv(u512 - pow(w512, 511, P)) % P, inv(u512 - 1), inv(useg - 1),
          inv(z - inv(root_of_unity(log_n)))]
    # periodic columns at z (Lagrange over the 512-th roots of unity)
    c = (pow(u512, 512, P) - 1) * inv(512) % P
    px = py = 0
    wr = 1
    for r in range(512):
        e, t = r >> 8, r & 255
        if t < 252:
            lr = c * wr % P * inv(u512 - wr) % P
            cx, cy = const_points[2 + 252 * e + t]
            px, py = (px + lr * cx) % P, (py + lr * cy) % P
        wr = wr * w512 % P
    acc = 0
    for l in range(LANES):
        X, Y, S, M, I = tz[5 * l:5 * l + 5]
        Xn, Yn, _s, Mn, _i = tzw[5 * l:5 * l + 5]
        a = alpha_pows[13 * l:13 * l + 13]
        bit = (M - 2 * Mn) % P
        nb = (1 - bit) % P
        c1 = bit * (bit - 1)
        c2 = bit * (S * (X - px) - (Y - py))
        c3 = bit * (S * S - X - px - Xn) + nb * (Xn - X)
        c4 = bit * (S * (X - Xn) - Y - Yn) + nb * (Yn - Y)
        acc += (a[0] * c1 + a[1] * c2 + a[2] * c3 + a[3] * c4) % P * iz[0]
        acc += a[4] * (I * (X - px) - 1) % P * iz[1] + a[5] * M % P * iz[2]
        acc += (a[6] * (Xn - X) + a[7] * (Yn - Y)) % P * iz[3] + a[8] * (Mn - X) % P * iz[4]
        acc += (a[9] * (X - shift[0]) + a[10] * (Y - shift[1])) % P * iz[5] + a[11] * (M - x0[l]) % P * iz[6]
        acc += a[12] * (X - outs[l]) % P * iz[7]
    return acc % P


def host_intt(vals, log_n):
    """natural-order inverse NTT of python ints (the <= 512-point last FRI layer)."""
    n = 1 << log_n
    a = list(vals)
    bits = log_n
    for i in range(n):
        r = int(format(i, "0%db" % bits)[::-1], 2) if bits else 0
        if r > i:
            a[i], a[r] = a[r], a[i]
    w = pow(root_of_unity(log_n), -1, P)
    h = 1
    while h < n:
        wh = pow(w, n // (2 * h), P)
        for b in range(0, n, 2 * h):
            t = 1
            for k in range(h):
                u, v = a[b + k], a[b + k + h] * t % P
                a[b + k], a[b + k + h] = (u + v) % P, (u - v) % P
                t = t * wh % P
        h *= 2
    ninv = pow(n, -1, P)
    return [x * ninv % P for x in a]


class ProofError(RuntimeError):
    pass


# ------------------------------------------------------------------ stage backend on libspg
class GpuBackend:
    """Stage calls on torch CUDA tensors (int64 limbs, last dimension 4) through the stage-level C-ABI."""

    def __init__(self, ctx):
        import torch
        self.torch, self.ctx, self.lib, self.h = torch, ctx, ctx._lib, ctx._h
        self.dev = torch.device("cuda", ctx.device)
        self._ticks = []
        # stage kernels and torch's collectives must be ordered on one stream
        with torch.cuda.device(self.dev):
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)

    def _chk(self, rc):
        self.ctx._check(rc)

    # ---- optional stage timing (device events on the shared stream + host clock), read by bench.py
    def tick(self, name):
        import time
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record(self.torch.cuda.current_stream())
        self._ticks.append((name, ev, time.perf_counter()))

    def reset_ticks(self):
        self._ticks = []

    def stage_times(self):
        """{stage: (device ms, host wall ms)} between consecutive ticks of the last proof"""
        self.torch.cuda.synchronize()
        out = {}
        for (n0, e0, t0), (_n1, e1, t1) in zip(self._ticks, self._ticks[1:]):
            d, w = out.get(n0, (0.0, 0.0))
            out[n0] = (d + e0.elapsed_time(e1), w + 1e3 * (t1 - t0))
        return out

    @staticmethod
    def _felts(values):
        arr = ints_to_limbs(values)
        return arr, arr.ctypes.data_as(C.c_void_p)

    def felts(self, *shape):
        return self.torch.empty(tuple(shape) + (4,), dtype=self.torch.int64, device=self.dev)

    def upload(self, np_limbs):
        return self.torch.from_numpy(np.ascontiguousarray(np_limbs).view(np.int64)).to(self.dev)

    def lde_coeffs(self, cols, log_n, n_cols, out, offset=None, mont=False):
        keep, op = self._felts([offset]) if offset is not None else (None, None)
        flags = SPG_DEVICE_PTRS | SPG_NO_SYNC | (SPG_MONT_OUT if mont else 0)      # stream-ordered with what follows
        self._chk(self.lib.spg_lde_coeffs(self.h, C.c_void_p(cols.data_ptr()), log_n, n_cols, op, C.c_void_p(out.data_ptr()), flags))

    def lde_cosets(self, coeffs, log_n, n_cols, first, count, out):
        self._chk(self.lib.spg_lde_cosets(self.h, C.c_void_p(coeffs.data_ptr()), log_n, n_cols, LOG_BLOWUP, first, count,
                                          C.c_void_p(out.data_ptr()), SPG_DEVICE_PTRS | SPG_NO_SYNC))

    def merkle(self, table, n_cols, rows, n_cosets):
        n_leaves = rows // 8 * n_cosets
        tree = self.torch.empty((2 * n_leaves - 1) * 32, dtype=self.torch.uint8, device=self.dev)
        self._chk(self.lib.spg_stage_merkle(self.h, C.c_void_p(table.data_ptr()), n_cols, rows, n_cosets, C.c_void_p(tree.data_ptr())))
        return tree

    def root(self, tree):
        """The sub-tree root as a 32-byte tensor on the backend's device (gathered with one small collective)."""
        return tree[-32:]

    def check_oods(self, log_n, chain_log, x0, outs, alpha, z, oods):
        """Prover self-check CP(z) == sum z^m H_m(z^4), host arithmetic inside libspg."""
        k1, p1 = self._felts(x0); k2, p2 = self._felts(outs); k3, p3 = self._felts([alpha]); k4, p4 = self._felts([z])
        k5, p5 = self._felts(oods)
        rc = self.lib.spg_stage_check_oods(self.h, log_n, chain_log, p1, p2, p3, p4, p5)
        if rc == -5:
            raise ProofError(self.lib.spg_last_error(self.h).decode())
        self._chk(rc)

    def table_bytes(self, t):
        """raw device representation of a small table (the last FRI layer), for the cross-rank gather"""
        return t.reshape(-1).cpu().numpy().tobytes()

    def bytes_tensor(self, b):
        """bytes -> uint8 tensor on the device the collectives run on"""
        return self.torch.frombuffer(bytearray(b), dtype=self.torch.uint8).to(self.dev)

    def last_layer(self, parts, log_rows_last, n_folds):
        """parts: table_bytes of every rank in rank order (= coset-major [8][n_last]).  Returns the serialised
        low coefficients; raises ProofError when the layer is not of low degree."""
        raw = np.frombuffer(b"".join(parts), dtype=np.uint64).copy()
        n_last = 1 << log_rows_last
        assert raw.size == 8 * n_last * 4
        out = np.empty(n_last * 32, dtype=np.uint8)
        rc = self.lib.spg_stage_last_layer(self.h, raw.ctypes.data_as(C.c_void_p), log_rows_last, n_folds,
                                           out.ctypes.data_as(C.c_void_p))
        if rc == -5:
            raise ProofError(self.lib.spg_last_error(self.h).decode())
        self._chk(rc)
        return out.tobytes()

    def air(self, t_lde, log_n, chain_log, first, jj0, n_even, x0, outs, alpha, cp):
        k1, p1 = self._felts(x0); k2, p2 = self._felts(outs); k3, p3 = self._felts([alpha])
        self._chk(self.lib.spg_stage_air(self.h, log_n, chain_log, C.c_void_p(t_lde.data_ptr()), first, jj0, n_even, p1, p2, p3,
                                         C.c_void_p(cp.data_ptr())))

    def cp_split(self, cp, log_n, jj0, n_even, hev):
        self._chk(self.lib.spg_stage_cp_split(self.h, log_n, C.c_void_p(cp.data_ptr()), jj0, n_even, C.c_void_p(hev.data_ptr())))

    def poly_eval(self, cols, pt_idx, pts, log_n):
        n = len(cols)
        ptrs = (C.c_void_p * n)(*[c.data_ptr() for c in cols])
        idx = (C.c_int * n)(*pt_idx)
        kp, pp = self._felts(pts)
        out = np.empty((n, 4), dtype=np.uint64)
        self._chk(self.lib.spg_stage_poly_eval(self.h, log_n, ptrs, idx, n, pp, len(pts), out.ctypes.data_as(C.c_void_p)))
        return limbs_to_ints(out)

    def deep(self, t_lde, h_lde, log_n, first, n_cosets, z, gamma, oods, out):
        scratch = self.felts(3, n_cosets, 1 << log_n)
        k1, p1 = self._felts([z]); k2, p2 = self._felts([gamma]); k3, p3 = self._felts(oods)
        self._chk(self.lib.spg_stage_deep(self.h, log_n, C.c_void_p(t_lde.data_ptr()), C.c_void_p(h_lde.data_ptr()), first, n_cosets,
                                          p1, p2, p3, C.c_void_p(scratch.data_ptr()), C.c_void_p(out.data_ptr())))
        self.ctx.synchronize()

    def fri_fold(self, layer, log_rows, first, n_cosets, beta, layer_index, out):
        k1, p1 = self._felts([beta])
        self._chk(self.lib.spg_stage_fri_fold(self.h, C.c_void_p(layer.data_ptr()), log_rows, first, n_cosets, p1, layer_index,
                                              C.c_void_p(out.data_ptr())))

    def open(self, table, n_cols, rows, n_cosets, tree, idx):
        """idx: LOCAL leaf indices.  Returns [(leaf_bytes, path_bytes)] per index."""
        if not idx:
            return []
        n_leaves = rows // 8 * n_cosets
        levels = n_leaves.bit_length() - 1
        ia = np.asarray(idx, dtype=np.uint32)
        lv = np.empty((len(idx), 8 * n_cols * 32), dtype=np.uint8)
        pa = np.empty((len(idx), max(levels, 1) * 32), dtype=np.uint8)
        self._chk(self.lib.spg_stage_open(self.h, C.c_void_p(table.data_ptr()), n_cols, rows, n_cosets, C.c_void_p(tree.data_ptr()),
                                          ia.ctypes.data_as(C.c_void_p), len(idx), lv.ctypes.data_as(C.c_void_p),
                                          pa.ctypes.data_as(C.c_void_p)))
        return [(lv[k].tobytes(), pa[k, :levels * 32].tobytes()) for k in range(len(idx))]

    def download_ints(self, t):
        """device table (Montgomery form) -> canonical python ints"""
        rinv = pow(R_MOD_P, -1, P)
        return [v * rinv % P for v in limbs_to_ints(t.reshape(-1, 4).cpu().numpy().view(np.uint64))]

    def const_points(self):
        import json
        import os
        prm = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "curve_params.json")))
        b = {k: (int(v[0], 16), int(v[1], 16)) for k, v in prm["BASE_POINTS"].items()}

        def dbl(pt):
            x, y = pt
            m = (3 * x * x + 1) * pow(2 * y, -1, P) % P
            nx = (m * m - 2 * x) % P
            return nx, (m * (x - nx) - y) % P
        pts = [b["SHIFT_POINT"], b["EC_GEN"]]
        for name, cnt in zip(("P0", "P1", "P2", "P3"), prm["CHAIN_LENGTHS"]):
            q = b[name]
            for _ in range(cnt):
                pts.append(q)
                q = dbl(q)
        return pts


class TorchComm:
    """The collectives the driver needs, on torch.distributed (NCCL for CUDA tensors, gloo on the CPU)."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def all_gather_into(self, out, inp):
        if self.world == 1:
            out.copy_(inp.reshape(out.shape))
            return
        import torch.distributed as dist
        dist.all_gather_into_tensor(out, inp)

    def broadcast(self, t, src):
        if self.world > 1:
            import torch.distributed as dist
            dist.broadcast(t, src)

    def all_gather_bytes(self, t):
        """t: a small uint8 tensor on the collective's device (same length on every rank) -> list of bytes per rank"""
        if self.world == 1:
            return [t.cpu().numpy().tobytes()]
        import torch
        import torch.distributed as dist
        out = torch.empty(self.world * t.numel(), dtype=torch.uint8, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous())
        raw = out.cpu().numpy().tobytes()
        k = t.numel()
        return [raw[r * k:(r + 1) * k] for r in range(self.world)]

    def sum_bytes(self, t):
        """element-wise sum over ranks of a uint8 tensor (used with disjoint non-zero regions: a scatter-free gather)"""
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy().tobytes()

    def all_gather_obj(self, obj):
        if self.world == 1:
            return [obj]
        import torch.distributed as dist
        out = [None] * self.world
        dist.all_gather_object(out, obj)
        return out


# ------------------------------------------------------------------ the sharded prover
def prove_sharded(be, comm, trace_cols_local, log_n, chain_log, x0, outs, n_queries=30, check=True, probe=None):
    """trace_cols_local: this rank's block of trace columns, tensor [my_cols][N][4] canonical (columns
    [rank*per, ...) with per = ceil(25/world)).  Returns the proof bytes on every rank.
    probe (tests): callable(stage_name, **tensors_and_challenges) invoked after every stage with the stage's device
    tables and the Fiat-Shamir challenges, so that parity tests can compare sampled values of the very run that
    produced the proof."""
    rank, world = comm.rank, comm.world
    assert BLOWUP % world == 0
    n, cs = 1 << log_n, BLOWUP // world
    first = rank * cs
    per = -(-N_COLS // world)
    c0, c1 = min(N_COLS, rank * per), min(N_COLS, (rank + 1) * per)
    my_cols = c1 - c0
    log_rows = fri_log_rows(log_n)
    n_folds = len(log_rows) - 1

    seed = (log_n.to_bytes(4, "little") + chain_log.to_bytes(4, "little") + n_queries.to_bytes(4, "little")
            + b"".join(ser(v) for v in x0) + b"".join(ser(v) for v in outs))
    ch = Channel(seed)
    tick = getattr(be, "tick", lambda name: None)
    probe = probe or (lambda name, **kw: None)
    if hasattr(be, "reset_ticks"):
        be.reset_ticks()
    proof = [b"SPGP", (1).to_bytes(4, "little"), log_n.to_bytes(4, "little"), chain_log.to_bytes(4, "little"),
             n_queries.to_bytes(4, "little"), n_folds.to_bytes(4, "little")] + [ser(v) for v in x0] + [ser(v) for v in outs]

    def commit(table, n_cols, rows):
        tree = be.merkle(table, n_cols, rows, cs)
        roots = comm.all_gather_bytes(be.root(tree))
        top = top_levels(roots)
        return tree, top

    # 1. interpolation (column-sharded) -> ONE all-gather -> coset evaluation (coset-sharded) + commitment
    tick("interpolate")
    mine = be.felts(per, n)
    if my_cols:
        be.lde_coeffs(trace_cols_local, log_n, my_cols, mine, mont=True)
    coefs = be.felts(per * world, n)
    tick("all_gather_coeffs")
    comm.all_gather_into(coefs, mine)
    tick("lde_trace")
    t_lde = be.felts(cs, N_COLS, n)
    be.lde_cosets(coefs, log_n, N_COLS, first, cs, t_lde)
    probe("lde_trace", coefs=coefs, t_lde=t_lde)
    tick("merkle_trace")
    tree_t, top_t = commit(t_lde, N_COLS, n)
    root_t = top_t[-1][0]
    ch.absorb(root_t)
    # 2. composition on the even cosets this rank owns; chunk split; chunk exchange; chunk LDE + commitment
    tick("air_composition")
    alpha = ch.draw_felt()
    even = [jj for jj in range(4) if first <= 2 * jj < first + cs]
    hev = be.felts(4, n)
    if even:
        cp = be.felts(len(even), n)
        be.air(t_lde, log_n, chain_log, first, even[0], len(even), x0, outs, alpha, cp)
        be.cp_split(cp, log_n, even[0], len(even), hev)
        probe("air_composition", alpha=alpha, cp=cp, even=even)
    hv = hev.view(4, n // 4, 4, 4)
    tick("chunk_exchange")
    if world > 1:
        for jj in range(4):
            piece = hv[:, :, jj, :].contiguous()
            comm.broadcast(piece, (2 * jj) // cs)
            hv[:, :, jj, :] = piece
    tick("lde_chunks")
    h_coef = be.felts(4, n)
    be.lde_coeffs(hev, log_n, 4, h_coef, offset=pow(GEN, -3, P), mont=False)
    h_lde = be.felts(cs, 4, n)
    be.lde_cosets(h_coef, log_n, 4, first, cs, h_lde)
    probe("lde_chunks", hev=hev, h_coef=h_coef, h_lde=h_lde)
    tick("merkle_chunks")
    tree_h, top_h = commit(h_lde, 4, n)
    root_h = top_h[-1][0]
    ch.absorb(root_h)
    # 3. out-of-domain values (every rank holds all coefficient columns)
    tick("oods_eval")
    z = ch.draw_felt()
    wn = root_of_unity(log_n)
    zw, z4 = z * wn % P, pow(z, 4, P)
    cols = [coefs[c] for c in range(N_COLS)] * 2 + [h_coef[m] for m in range(4)]
    pidx = [0] * N_COLS + [1] * N_COLS + [2] * 4
    # the 54 evaluations are dealt round-robin to the ranks (every rank holds every coefficient column) and the
    # 32-byte results gathered
    mine_items = list(range(rank, len(cols), world))
    vals = be.poly_eval([cols[k] for k in mine_items], [pidx[k] for k in mine_items], [z, zw, z4], log_n) if mine_items else []
    per_rank = -(-len(cols) // world)
    buf = b"".join(v.to_bytes(32, "little") for v in vals).ljust(32 * per_rank, b"\0")
    oods = [0] * len(cols)
    for r, part in enumerate(comm.all_gather_bytes(be.bytes_tensor(buf))):
        for i, k in enumerate(range(r, len(cols), world)):
            oods[k] = int.from_bytes(part[32 * i:32 * i + 32], "little")
    if check:
        be.check_oods(log_n, chain_log, x0, outs, alpha, z, oods)
    probe("oods_eval", z=z, oods=oods)
    ob = b"".join(ser(v) for v in oods)
    ch.absorb(ob)
    # 4. DEEP quotient on the local cosets
    tick("deep_quotient")
    gamma = ch.draw_felt()
    layers, trees = [be.felts(cs, n)], [None]
    be.deep(t_lde, h_lde, log_n, first, cs, z, gamma, oods, layers[0])
    probe("deep_quotient", gamma=gamma, layer0=layers[0])
    # 5. FRI.  Only the first fold works on sharded data: its output (N/8 rows per coset, 32 MB at 2^20) is
    # all-gathered once, and every rank then folds and commits the remaining, geometrically shrinking layers
    # locally -- no further collective or root exchange on the latency-bound tail of the protocol.
    tick("fri")
    fri_roots = []
    full_tops = [None]
    for l in range(1, n_folds + 1):
        beta = ch.draw_felt()
        rows_l = 1 << log_rows[l]
        if l == 1:
            part = be.felts(cs, rows_l)
            be.fri_fold(layers[0], log_rows[0], first, cs, beta, 0, part)
            if world > 1:
                nxt = be.felts(BLOWUP, rows_l)
                comm.all_gather_into(nxt, part)                   # rank r holds cosets [r cs, (r+1) cs): coset-major
            else:
                nxt = part
        else:
            nxt = be.felts(BLOWUP, rows_l)
            be.fri_fold(layers[l - 1], log_rows[l - 1], 0, BLOWUP, beta, l - 1, nxt)
        probe("fri", l=l, beta=beta, prev=layers[l - 1], layer=nxt)
        tree = be.merkle(nxt, 1, rows_l, BLOWUP)
        root = be.root(tree).cpu().numpy().tobytes()
        layers.append(nxt); trees.append(tree); full_tops.append([[root]])
        fri_roots.append(root)
        ch.absorb(root)
    lb = be.last_layer([be.table_bytes(layers[-1])], log_rows[-1], n_folds)
    ch.absorb(lb)
    proof += [root_t, root_h, ob] + fri_roots + [lb]
    # 6. queries: every rank opens the leaves that live in its cosets
    tick("queries")
    # (table, columns, rows, tree, top levels, first coset held, cosets held); FRI layers are replicated (opened by rank 0)
    tables = [(t_lde, N_COLS, n, tree_t, top_t, first, cs), (h_lde, 4, n, tree_h, top_h, first, cs)]
    tables += [(layers[l], 1, 1 << log_rows[l], trees[l], full_tops[l], 0, BLOWUP) for l in range(1, n_folds + 1)]
    qidx = []                              # per query, per table: (j, i')
    for _ in range(n_queries):
        idx = ch.draw_index(n)
        j, ip = idx // (n // 8), idx % (n // 8)
        row = [(j, ip), (j, ip)]
        for l in range(1, n_folds + 1):
            ip %= (1 << log_rows[l]) // 8
            row.append((j, ip))
        qidx.append(row)
    # every (query, table) opening has one owner and a size known to all ranks: each rank writes its openings into a
    # zeroed buffer with a fixed layout and the buffers are summed (one small all-reduce instead of pickled objects)
    sizes = []
    for (_table, n_cols, rows, _tree, _top, _tf, tcs) in tables:
        levels = (rows // 8 * tcs).bit_length() - 1
        sizes.append((8 * n_cols * 32, levels * 32))
    per_query = sum(a + b for a, b in sizes)
    t_off = [sum(a + b for a, b in sizes[:t]) for t in range(len(tables))]
    buf = bytearray(per_query * n_queries)
    for t, (table, n_cols, rows, tree, _top, tfirst, tcs) in enumerate(tables):
        if tcs == BLOWUP and world > 1 and rank != 0:
            continue                       # replicated table: one opener is enough
        want = [(q, (j - tfirst) * (rows // 8) + ip) for q, rowq in enumerate(qidx) for (j, ip) in [rowq[t]]
                if tfirst <= j < tfirst + tcs]
        res = be.open(table, n_cols, rows, tcs, tree, [w[1] for w in want])
        for (q, _li), (leaf, path) in zip(want, res):
            o = q * per_query + t_off[t]
            buf[o:o + len(leaf)] = leaf
            buf[o + sizes[t][0]:o + sizes[t][0] + len(path)] = path
    allb = comm.sum_bytes(be.bytes_tensor(bytes(buf)))
    for q in range(n_queries):
        for t, (_table, _nc, _rows, _tree, top, _tf, tcs) in enumerate(tables):
            o = q * per_query + t_off[t]
            owner = qidx[q][t][0] // tcs
            proof += [allb[o:o + sizes[t][0]], allb[o + sizes[t][0]:o + sizes[t][0] + sizes[t][1]]] + top_path(top, owner)
    tick("end")
    return b"".join(proof)


class Prover:
    def __init__(self, ctx, rank=0, world=1, backend=None, comm=None):
        self.ctx, self.rank, self.world = ctx, rank, world
        self.comm = comm or TorchComm(rank, world)
        self._be = backend

    @property
    def be(self):
        if self._be is None:
            self._be = GpuBackend(self.ctx)
        return self._be

    # ---- witness ---------------------------------------------------------------------------------
    def witness(self, log_n, chain_log, x0, ys):
        """ys: list of 5 lists of ints (or a (5 * inst, 4) uint64 array).  Returns the (25 N, 4) trace."""
        if not isinstance(ys, np.ndarray):
            ys = ints_to_limbs([v for lane in ys for v in lane])
        return self.ctx.pedersen_chain_trace(log_n, chain_log, x0, ys)

    # ---- single GPU: everything inside libspg ------------------------------------------------------
    def prove_host(self, trace, log_n, chain_log, x0, n_queries=30):
        if self.world > 1:
            cols, outs = self.shard_host_trace(trace, log_n)
            return prove_sharded(self.be, self.comm, cols, log_n, chain_log, x0, outs, n_queries)
        return self.ctx.prove(trace, log_n, chain_log, x0, n_queries)

    def prove_device(self, trace_ptr, log_n, chain_log, x0, n_queries=30):
        assert self.world == 1, "multi-GPU: use prove_sharded_device with this rank's column block"
        return self.ctx.prove(None, log_n, chain_log, x0, n_queries, device_ptr=trace_ptr)

    # ---- multi GPU ---------------------------------------------------------------------------------
    def column_block(self):
        per = -(-N_COLS // self.world)
        return min(N_COLS, self.rank * per), min(N_COLS, (self.rank + 1) * per)

    def shard_host_trace(self, trace, log_n):
        """Upload only this rank's columns of a host trace ((25 N, 4) uint64); returns (device block, outs)."""
        n = 1 << log_n
        tr = np.ascontiguousarray(trace, dtype=np.uint64).reshape(N_COLS, n, 4)
        outs = limbs_to_ints(tr[[5 * l for l in range(LANES)], n - 1])
        c0, c1 = self.column_block()
        block = self.be.upload(tr[c0:c1]) if c1 > c0 else self.be.felts(1, n)
        return block, outs

    def prove_sharded_device(self, cols_local, log_n, chain_log, x0, outs, n_queries=30, probe=None):
        return prove_sharded(self.be, self.comm, cols_local, log_n, chain_log, x0, outs, n_queries, probe=probe)

    # ---- multi GPU, sequenced inside libspg (csrc/sharded.cu: NCCL called from C++, pipelined column exchange) --------
    def setup_comm(self):
        """Join libspg's own NCCL communicator: rank 0 creates the id, torch.distributed carries it to the others."""
        if getattr(self, "_comm_ready", False):
            return
        uid = None
        if self.world > 1:
            import torch.distributed as dist
            box = [self.ctx.comm_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            uid = box[0]
        self.ctx.comm_init(self.rank, self.world, uid)
        self._comm_ready = True

    def cyclic_columns(self):
        """the trace columns this rank owns in the C++ sharded prover: rank, rank + world, ..."""
        return list(range(self.rank, N_COLS, self.world))

    def prove_cyclic(self, cols_local, log_n, chain_log, x0, outs, n_queries=30, device_ptr=None):
        """cols_local: this rank's cyclic columns, (my_cols * N, 4) uint64 host array (or device_ptr).  Collective."""
        self.setup_comm()
        return self.ctx.prove_sharded(cols_local, log_n, chain_log, x0, outs, n_queries, device_ptr=device_ptr)

    def parallelism(self):
        if self.world == 1:
            return "1 GPU"
        return ("%d GPUs: columns dealt cyclically -> pipelined NCCL all-gather of coefficient columns -> %d coset(s) per GPU "
                "(sequenced in libspg, csrc/sharded.cu)" % (self.world, BLOWUP // self.world))
