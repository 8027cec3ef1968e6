"""Sampled-point parity of every prover stage of ONE run (TEST INFRASTRUCTURE).

`StageChecker` is handed to prover.prove_sharded as its `probe`: after each stage it reads the stage's device table
at a fixed sample of points and compares with the oracle applied to independently computed inputs:

  trace LDE      all 25 columns at the sample (rows i and i+1)         vs  the C oracle's LDE of the host trace
  composition    CP at the sample's even-coset points                  vs  oracle/stark.py Air.composition on the C-LDE values
  chunk split    H_m on g^4 <w_N> at sampled positions                 vs  the size-4 inverse DFT of the run's CP values
  chunk LDE      all 4 columns at the sample                           vs  the C oracle's LDE of the run's chunk values
  out-of-domain  all 54 values                                         vs  Horner on C-oracle coefficients
  DEEP quotient  at the sample (whole fold groups)                     vs  oracle/stark.py deep_quotient on C-LDE values
  FRI folds      every layer at sampled groups (all groups when small) vs  oracle/stark.py fold8 on the previous layer

With `n_groups` = 512 the sample is 4096 points (SURVEY.md section 8(d) cfg-3a asks for 4096 sampled points at 2^20).
The sample is made of whole fold groups {i' + k N/8} of a coset, so the DEEP values checked are exactly the inputs of
the sampled first-layer folds.  Works on any stage backend (GpuBackend on the GPU, CpuBackend under pytest on CPU).
"""
import random

import numpy as np

from oracle import clib, stark
from oracle.params import FIELD_PRIME as P, root_of_unity
from stark_perpetual_b200._lib import ints_to_limbs, limbs_to_ints

GEN = 3


def _inv(a):
    return pow(a % P, -1, P)


class StageChecker:
    def __init__(self, be, trace_host, log_n, chain_log, x0, outs, n_groups=512, seed=2020):
        import torch
        self.torch, self.be = torch, be
        self.log_n, self.n, self.chain_log, self.x0, self.outs = log_n, 1 << log_n, chain_log, x0, outs
        n = self.n
        rng = random.Random(seed)
        grp = n // 8
        n_groups = min(n_groups, 2 * grp)
        # groups (j, i'): the first half forced onto even cosets so that the composition has a full sample too
        keys = set()
        while len(keys) < n_groups:
            j = rng.randrange(8)
            if len(keys) < n_groups // 2:
                j &= 6
            keys.add((j, rng.randrange(grp)))
        self.groups = sorted(keys)
        self.points = [(j, ip + k * grp) for (j, ip) in self.groups for k in range(8)]
        self.trace_host = np.ascontiguousarray(trace_host, dtype=np.uint64).reshape(25, n, 4)
        self.seen = []
        self._oracle_trace_lde()

    # ---- helpers
    def x_of(self, j, i, log_rows=None, g=GEN):
        lr = self.log_n if log_rows is None else log_rows
        return g * pow(root_of_unity(lr + 3), j + 8 * i, P) % P

    def sample(self, table, pts, n_cols):
        """table: backend tensor [cosets][n_cols][rows][4] (or [cosets][rows][4] when n_cols is None) holding all 8 cosets;
        returns per point the list of column values (canonical ints)."""
        t = self.torch
        jt = t.tensor([p[0] for p in pts], dtype=t.long, device=table.device)
        it = t.tensor([p[1] for p in pts], dtype=t.long, device=table.device)
        if n_cols is None:
            vals = self.be.download_ints(table[jt, it])
            return vals
        sel = table[jt, :, it]                                   # [S][n_cols][4]
        vals = self.be.download_ints(sel)
        return [vals[k * n_cols:(k + 1) * n_cols] for k in range(len(pts))]

    def _oracle_trace_lde(self):
        """C-oracle LDE of the host trace, kept only at the sampled points (rows i and i + 1)."""
        n, log_n = self.n, self.log_n
        clib.use_all_cores()
        want = sorted({(j, i) for (j, i) in self.points} | {(j, (i + 1) % n) for (j, i) in self.points})
        self.t_vals = {pt: [0] * 25 for pt in want}
        jj = np.array([p[0] for p in want]); ii = np.array([p[1] for p in want])
        for c0 in range(0, 25, 13):              # two batches: at 2^20 one batch of extended columns is 3.5 GB of host memory
            step = min(13, 25 - c0)
            cols = self.trace_host[c0:c0 + step].reshape(-1, 4)
            lde = clib.lde(cols, log_n, step, 3).reshape(8, step, n, 4)
            for c in range(step):
                vals = limbs_to_ints(np.ascontiguousarray(lde[jj, c, ii]))
                for pt, v in zip(want, vals):
                    self.t_vals[pt][c0 + c] = v
            del lde

    # ---- the probe
    def __call__(self, name, **kw):
        getattr(self, "on_" + name)(**kw)
        self.seen.append(name)

    def on_lde_trace(self, coefs, t_lde):
        n = self.n
        pts = self.points + [(j, (i + 1) % n) for (j, i) in self.points]
        got = self.sample(t_lde, pts, 25)
        for pt, g in zip(pts, got):
            assert g == self.t_vals[pt], ("trace LDE", pt)

    def on_air_composition(self, alpha, cp, even):
        assert even == [0, 1, 2, 3]
        n, log_n = self.n, self.log_n
        self.alpha = alpha
        apows = [pow(alpha, k, P) for k in range(stark.LANES * stark.N_CONSTRAINTS)]
        air = stark.Air(log_n, self.chain_log, self.x0, self.outs)
        pts = [(j, i) for (j, i) in self.points if j % 2 == 0]
        got = self.sample(cp, [(j // 2, i) for (j, i) in pts], None)
        assert len(pts) >= len(self.points) // 2 - 8
        for (j, i), g in zip(pts, got):
            x = self.x_of(j, i)
            px, py = air.periodic_at(x)
            want = air.composition(self.t_vals[(j, i)], self.t_vals[(j, (i + 1) % n)], px, py, air.inv_zerofiers(x), apows)
            assert g == want, ("composition", j, i)
        # inputs of the sampled chunk-split positions, read from the run's own CP
        q = n // 4
        rng = random.Random(7)
        self.split_pos = [(rng.randrange(4), rng.randrange(q)) for _ in range(min(1024, q))]
        quad = [(jj, ip + k * q) for (jj, ip) in self.split_pos for k in range(4)]
        vals = self.sample(cp, quad, None)
        self.split_in = [vals[4 * k:4 * k + 4] for k in range(len(self.split_pos))]

    def on_lde_chunks(self, hev, h_coef, h_lde):
        n, log_n = self.n, self.log_n
        t = self.torch
        # chunk split at the sampled positions: H_m(x^4) x^m = 1/4 sum_k i^(-m k) CP(x i^k)
        iota_inv, inv4 = _inv(root_of_unity(2)), _inv(4)
        pos = t.tensor([jj + 4 * ip for (jj, ip) in self.split_pos], dtype=t.long, device=hev.device)
        got = self.be.download_ints(hev[:, pos])                           # [4][S]
        S = len(self.split_pos)
        for s, ((jj, ip), vals) in enumerate(zip(self.split_pos, self.split_in)):
            xi = _inv(self.x_of(2 * jj, ip))
            for m in range(4):
                want = sum(pow(iota_inv, m * k, P) * vals[k] for k in range(4)) % P * inv4 % P * pow(xi, m, P) % P
                assert got[m * S + s] == want, ("chunk split", jj, ip, m)
        # chunk LDE: C-oracle LDE of the run's chunk values (they live on g^4 <w_N>: offset g^-3)
        hv = ints_to_limbs(self.be.download_ints(hev))
        self.hev_host = hv
        lde = clib.lde(hv, log_n, 4, 3, offset=pow(GEN, -3, P)).reshape(8, 4, n, 4)
        jj = np.array([p[0] for p in self.points]); ii = np.array([p[1] for p in self.points])
        self.h_vals = {}
        cols = [limbs_to_ints(np.ascontiguousarray(lde[jj, m, ii])) for m in range(4)]
        for k, pt in enumerate(self.points):
            self.h_vals[pt] = [cols[m][k] for m in range(4)]
        got = self.sample(h_lde, self.points, 4)
        for pt, g in zip(self.points, got):
            assert g == self.h_vals[pt], ("chunk LDE", pt)

    def on_oods_eval(self, z, oods):
        n, log_n = self.n, self.log_n
        self.z, self.oods = z, oods
        zw = z * root_of_unity(log_n) % P
        t_coef = clib.ntt(self.trace_host.reshape(-1, 4), log_n, inverse=True, order=2)
        want = clib.poly_eval(t_coef, n, [z] * 25) + clib.poly_eval(t_coef, n, [zw] * 25)
        # chunk m: values on g^4 <w_N>  =>  H_m(y) = sum_k intt_k (y / g^4)^k
        h_coef = clib.ntt(self.hev_host, log_n, inverse=True, order=2)
        want += clib.poly_eval(h_coef, n, [pow(z * _inv(GEN), 4, P)] * 4)
        assert oods == want, "out-of-domain values"

    def on_deep_quotient(self, gamma, layer0):
        z, log_n = self.z, self.log_n
        zw, z4 = z * root_of_unity(log_n) % P, pow(z, 4, P)
        gp = [pow(gamma, k, P) for k in range(54)]
        got = self.sample(layer0, self.points, None)
        self.deep_vals = {}
        for pt, g in zip(self.points, got):
            want = stark.deep_quotient(self.t_vals[pt], self.h_vals[pt], self.x_of(*pt), z, zw, z4, self.oods, gp)
            assert g == want, ("DEEP quotient", pt)
            self.deep_vals[pt] = want

    def on_fri(self, l, beta, prev, layer):
        log_rows = stark_log_rows(self.log_n)
        lr_in, rows_out = log_rows[l - 1], 1 << log_rows[l]
        g_l = pow(GEN, 8 ** (l - 1), P)
        if l == 1:
            groups = self.groups
            ins = [[self.deep_vals[(j, ip + k * rows_out)] for k in range(8)] for (j, ip) in groups]
        else:
            rng = random.Random(100 + l)
            groups = [(j, ip) for j in range(8) for ip in range(rows_out)]
            if len(groups) > 1024:
                groups = rng.sample(groups, 1024)
            vals = self.sample(prev, [(j, ip + k * rows_out) for (j, ip) in groups for k in range(8)], None)
            ins = [vals[8 * k:8 * k + 8] for k in range(len(groups))]
        got = self.sample(layer, groups, None)
        for (j, ip), vin, g in zip(groups, ins, got):
            want = stark.fold8(vin, self.x_of(j, ip, lr_in, g_l), beta)
            assert g == want, ("FRI fold", l, j, ip)

    def assert_complete(self):
        want = ["lde_trace", "air_composition", "lde_chunks", "oods_eval", "deep_quotient", "fri"]
        assert [s for s in want if s not in self.seen] == [], self.seen


def stark_log_rows(log_n):
    lr = [log_n]
    while (1 << lr[-1]) > stark.LAST_LAYER_MAX:
        lr.append(lr[-1] - 3)
    return lr
