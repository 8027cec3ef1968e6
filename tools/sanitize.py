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
    # round-2 kernels: the C++ sharded prover with a one-rank communicator, math_utils ops, state trees, hash chain
    tr10 = pv.witness(10, 1, x0, ys)
    outs = limbs_to_ints(tr10.reshape(25, 1024, 4)[[5 * l for l in range(5)], 1023])
    ctx.comm_init(0, 1)
    assert ctx.prove_sharded(tr10, 10, 1, x0, outs, 8) == proof
    if big:
        one = rand_felts(1 << 20, 9)
        assert np.array_equal(ctx.lde(one, 20, 1, 3)[:1 << 20], clib.lde(one, 20, 1, 3)[:1 << 20])     # k_ntt_tile, both passes, TMA load
    g = [int(v, 16) for v in ("1ef15c18599971b7beced415a40f0c7deacfd9b0d1819e03d723d8bc943cfca", "5668060aa49730b7be4801df46ec62de53ecd11abe43a32873000c36e8dc1f")]
    xy = ints_to_limbs(g).reshape(1, 8)
    o2, st2 = ctx.ec_op(2, np.tile(xy, (3, 1)), ints_to_limbs([5, 2**200 + 3, 1]))
    assert not st2.any() and np.array_equal(o2[2], xy[0])
    pk = ints_to_limbs([7, 9])
    h, st3 = ctx.position_hash(pk, np.array([1, -2], dtype=np.int64), np.array([0, 1, 1], dtype=np.uint64),
                               np.array([[5, 0]], dtype=np.uint64), np.array([3], dtype=np.int64), np.array([-4], dtype=np.int64))
    assert not st3.any()
    sib = ctx.merkle_update_siblings(6, [3, 9, 10])
    pr, nr, st4, _ = ctx.merkle_multi_update(6, [3, 9, 10], rand_felts(3, 1), rand_felts(3, 2), rand_felts(len(sib), 3))
    assert st4 == 0
    oc, st5 = ctx.hash_chain_rfold(rand_felts(6, 4), 3)
    assert not st5.any()
    # second AIR: witness kernels, public columns, constraint kernel, proof (2 signatures, 2^9 rows)
    from oracle import stark_ecdsa as se
    sigs = se.make_signatures(2, 5)
    ins = [ints_to_limbs(v) for v in ([s_[0] for s_ in sigs], [s_[1] for s_ in sigs], [s_[2] for s_ in sigs],
                                      [s_[3][0] for s_ in sigs], [s_[3][1] for s_ in sigs])]
    tr9 = ctx.ecdsa_air_trace(9, *ins)
    ep = ctx.prove_ecdsa(tr9, 9, ins[0], ins[3], 8)
    assert ostark.verify(ep, min_queries=8)["air"] == "ecdsa"
    print("sanitize workload OK")


if __name__ == "__main__":
    main()
