#!/usr/bin/env python3
"""Workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once, at sizes that exercise
the window-local warp barriers of the NTT (whole-workspace passes, both tile shapes), checked against the oracle.
    compute-sanitizer --tool racecheck python tools/sanitize.py"""
import os
import random
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import stark_perpetual_b200 as spg  # noqa: E402
from stark_perpetual_b200._lib import NTT_NAT_TO_REV, NTT_REV_TO_NAT, ints_to_limbs, limbs_to_ints  # noqa: E402
from stark_perpetual_b200.prover import Prover  # noqa: E402
from conftest import rand_felts  # noqa: E402
from oracle import clib, stark as ostark  # noqa: E402
from oracle.params import FIELD_PRIME as P  # noqa: E402
from oracle.pedersen import pedersen_hash  # noqa: E402


def main():
    big = "--big" in sys.argv
    ctx = spg.get_context(0)
    for log_n in (10, 11, 13) + ((20, 21) if big else ()):
        x = rand_felts(1 << log_n, log_n)
        f = ctx.ntt(x, log_n, False, NTT_NAT_TO_REV)
        assert np.array_equal(f, clib.ntt(x, log_n, False, 0)), log_n
        assert np.array_equal(ctx.ntt(f, log_n, True, NTT_REV_TO_NAT), x), log_n
    tr = rand_felts(2 << 11, 5)
    assert np.array_equal(ctx.lde(tr, 11, 2, 3), clib.lde(tr, 11, 2, 3))
    rng = random.Random(3)
    x0 = [rng.randrange(P) for _ in range(5)]
    ys = [[rng.randrange(P) for _ in range(2)] for _ in range(5)]
    pv = Prover(ctx)
    proof = pv.prove_host(pv.witness(10, 1, x0, ys), 10, 1, x0, n_queries=8)
    ostark.verify(proof, min_queries=8)
    out, st = ctx.pedersen_hash2(ints_to_limbs([1, 2]), ints_to_limbs([3, 4]))
    assert limbs_to_ints(out) == [pedersen_hash(1, 3), pedersen_hash(2, 4)]
    print("sanitize workload OK")


if __name__ == "__main__":
    main()
