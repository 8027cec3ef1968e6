"""CPU stand-in for stark_perpetual_b200.prover.GpuBackend, built on the oracle (TEST INFRASTRUCTURE).

It lets the multi-process driver `prove_sharded` -- column/coset sharding, the all-gather of coefficient
columns, the chunk exchange, sub-tree roots, distributed query openings -- run under gloo on CPU tensors, so
the N>1 host logic is exercised without GPUs.  Tables hold canonical values; only the driver-visible
artefacts (tree bytes, leaf bytes) follow the wire format.
"""
import numpy as np
import torch

from oracle import ntt as ontt, stark
from oracle.params import CONSTANT_POINTS, FIELD_PRIME as P, root_of_unity
from stark_perpetual_b200._lib import ints_to_limbs, limbs_to_ints

GEN = 3


class CpuBackend:
    def felts(self, *shape):
        return torch.zeros(tuple(shape) + (4,), dtype=torch.int64)

    def upload(self, np_limbs):
        return torch.from_numpy(np.ascontiguousarray(np_limbs).view(np.int64)).clone()

    @staticmethod
    def _get(t):
        return limbs_to_ints(t.contiguous().numpy().view(np.uint64).reshape(-1, 4))

    @staticmethod
    def _put(t, values):
        t.copy_(torch.from_numpy(ints_to_limbs(values).view(np.int64)).reshape(t.shape))

    def lde_coeffs(self, cols, log_n, n_cols, out, offset=None, mont=False):
        n = 1 << log_n
        vals = self._get(cols[:n_cols])
        a = 1 if offset is None else GEN * pow(offset, -1, P) % P        # coset the values live on
        ai = pow(a, -1, P)
        res = []
        for c in range(n_cols):
            cf = ontt.ntt(vals[c * n:(c + 1) * n], inverse=True)
            res += [v * pow(ai, k, P) % P for k, v in enumerate(cf)]
        self._put(out[:n_cols], res)

    def lde_cosets(self, coeffs, log_n, n_cols, first, count, out):
        n = 1 << log_n
        cf = self._get(coeffs[:n_cols])
        res = [None] * count
        for c in range(n_cols):
            cosets = ontt.lde(ontt.ntt(cf[c * n:(c + 1) * n]), 3, GEN)
            for jl in range(count):
                res[jl] = (res[jl] or []) + cosets[first + jl]
        self._put(out, [v for jl in range(count) for v in res[jl]])

    def _table(self, table, n_cols, rows, n_cosets):
        v = self._get(table)
        return [[v[(jl * n_cols + c) * rows:(jl * n_cols + c + 1) * rows] for c in range(n_cols)] for jl in range(n_cosets)]

    def merkle(self, table, n_cols, rows, n_cosets):
        tab = self._table(table, n_cols, rows, n_cosets)
        g = rows // 8
        leaves = [b"".join(stark.ser(col[ip + k * g]) for k in range(8) for col in tab[jl]) for jl in range(n_cosets)
                  for ip in range(g)]
        levels = stark.merkle_levels([stark.H(x) for x in leaves])
        return {"levels": levels, "leaves": leaves}

    def root(self, tree):
        return torch.from_numpy(np.frombuffer(tree["levels"][-1][0], dtype=np.uint8).copy())

    def check_oods(self, log_n, chain_log, x0, outs, alpha, z, oods):
        from stark_perpetual_b200 import prover
        apows = [pow(alpha, k, P) for k in range(65)]
        lhs = prover.composition_at(log_n, chain_log, x0, outs, apows, z, oods[:25], oods[25:50], CONSTANT_POINTS,
                                    CONSTANT_POINTS[0])
        if lhs != sum(pow(z, m, P) * oods[50 + m] for m in range(4)) % P:
            raise prover.ProofError("trace does not satisfy the AIR (composition mismatch at the out-of-domain point)")

    def table_bytes(self, t):
        return b"".join(v.to_bytes(32, "little") for v in self._get(t))

    def bytes_tensor(self, b):
        return torch.frombuffer(bytearray(b), dtype=torch.uint8)

    def last_layer(self, parts, log_rows_last, n_folds):
        from stark_perpetual_b200 import prover
        n_last = 1 << log_rows_last
        raw = b"".join(parts)
        vals = [int.from_bytes(raw[32 * k:32 * k + 32], "little") for k in range(8 * n_last)]     # [8][n_last]
        flat = [0] * (8 * n_last)
        for j in range(8):
            for i in range(n_last):
                flat[j + 8 * i] = vals[j * n_last + i]
        lc = prover.host_intt(flat, log_rows_last + 3)
        gli = pow(pow(GEN, 8 ** n_folds, P), -1, P)
        lc = [c * pow(gli, k, P) % P for k, c in enumerate(lc)]
        if any(lc[n_last:]):
            raise prover.ProofError("trace does not satisfy the AIR (FRI last layer is not of low degree)")
        return b"".join(stark.ser(v) for v in lc[:n_last])

    def air(self, t_lde, log_n, chain_log, first, jj0, n_even, x0, outs, alpha, cp):
        n = 1 << log_n
        air = stark.Air(log_n, chain_log, x0, outs)
        apows = [pow(alpha, k, P) for k in range(65)]
        px512, py512 = stark.periodic_points()
        g512 = pow(GEN, n // 512, P)
        px_lde, py_lde = ontt.lde(px512, 3, g512), ontt.lde(py512, 3, g512)
        cs = t_lde.shape[0]
        tab = self._table(t_lde, 25, n, cs)
        res = []
        for e in range(n_even):
            j = 2 * (jj0 + e)
            for i in range(n):
                x = stark.lde_point(log_n, j, i)
                cur = [tab[j - first][c][i] for c in range(25)]
                nxt = [tab[j - first][c][(i + 1) % n] for c in range(25)]
                res.append(air.composition(cur, nxt, px_lde[j][i % 512], py_lde[j][i % 512], air.inv_zerofiers(x), apows))
        self._put(cp, res)

    def cp_split(self, cp, log_n, jj0, n_even, hev):
        n, q = 1 << log_n, (1 << log_n) // 4
        v = self._get(cp)
        h = self._get(hev)
        iota_inv = pow(root_of_unity(2), -1, P)
        inv4 = pow(4, -1, P)
        for e in range(n_even):
            jj = jj0 + e
            for ip in range(q):
                x = stark.lde_point(log_n, 2 * jj, ip)
                xi = pow(x, -1, P)
                vals = [v[e * n + ip + k * q] for k in range(4)]
                for m in range(4):
                    s = sum(pow(iota_inv, m * k, P) * vals[k] for k in range(4)) % P
                    h[m * n + jj + 4 * ip] = s * inv4 % P * pow(xi, m, P) % P
        self._put(hev, h)

    def poly_eval(self, cols, pt_idx, pts, log_n):
        out = []
        for col, pi in zip(cols, pt_idx):
            acc = 0
            for c in reversed(self._get(col)):
                acc = (acc * pts[pi] + c) % P
            out.append(acc)
        return out

    def deep(self, t_lde, h_lde, log_n, first, n_cosets, z, gamma, oods, out):
        n = 1 << log_n
        tt, hh = self._table(t_lde, 25, n, n_cosets), self._table(h_lde, 4, n, n_cosets)
        gp = [pow(gamma, k, P) for k in range(54)]
        zw, z4 = z * root_of_unity(log_n) % P, pow(z, 4, P)
        res = [stark.deep_quotient([tt[jl][c][i] for c in range(25)], [hh[jl][m][i] for m in range(4)],
                                   stark.lde_point(log_n, first + jl, i), z, zw, z4, oods, gp)
               for jl in range(n_cosets) for i in range(n)]
        self._put(out, res)

    def fri_fold(self, layer, log_rows, first, n_cosets, beta, layer_index, out):
        rows, grp = 1 << log_rows, (1 << log_rows) // 8
        v = self._get(layer)
        g_l = pow(GEN, 8 ** layer_index, P)
        w = root_of_unity(log_rows + 3)
        res = [stark.fold8([v[jl * rows + ip + k * grp] for k in range(8)], g_l * pow(w, first + jl + 8 * ip, P) % P, beta)
               for jl in range(n_cosets) for ip in range(grp)]
        self._put(out, res)

    def open(self, table, n_cols, rows, n_cosets, tree, idx):
        return [(tree["leaves"][i], b"".join(stark.merkle_path(tree["levels"], i))) for i in idx]

    def download_ints(self, t):
        return self._get(t)

    def const_points(self):
        return CONSTANT_POINTS
