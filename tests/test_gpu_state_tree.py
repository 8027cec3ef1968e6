"""GPU parity: spg_position_hash_batch and spg_merkle_multi_update against oracle/state_tree.py
(restating position/hash.cairo:22-74, state/state.cairo:143-173, starkware/python/merkle_tree.py:4-44) and against the dense
Pedersen tree kernel."""
import random
import time

import numpy as np
import pytest

from conftest import rand_felts
from oracle import state_tree
from oracle.params import FIELD_PRIME as P
from oracle.pedersen import pedersen_hash
from stark_perpetual_b200._lib import ints_to_limbs, limbs_to_ints

pytestmark = pytest.mark.gpu


def _positions(rng, n, max_assets=6):
    pos = []
    for _ in range(n):
        k = rng.randrange(max_assets + 1)
        ids = sorted({rng.randrange(2**120) for _ in range(k)})
        assets = [(a, rng.randrange(-2**63, 2**63), rng.randrange(-2**63, 2**63)) for a in ids]
        pos.append((rng.randrange(P), rng.randrange(-2**63, 2**63), assets))
    return pos


def _csr(pos):
    off = np.cumsum([0] + [len(p[2]) for p in pos]).astype(np.uint64)
    flat = [a for p in pos for a in p[2]]
    aid = np.array([[a[0] & (2**64 - 1), a[0] >> 64] for a in flat], dtype=np.uint64).reshape(-1, 2)
    bal = np.array([a[1] for a in flat], dtype=np.int64)
    fi = np.array([a[2] for a in flat], dtype=np.int64)
    return ints_to_limbs([p[0] for p in pos]), np.array([p[1] for p in pos], dtype=np.int64), off, aid, bal, fi


def test_position_hash_vs_oracle(ctx):
    rng = random.Random(2024)
    pos = _positions(rng, 40)
    pos.append((1, 0, []))                                                         # empty position
    pos.append((P - 1, 2**63 - 1, [(2**120 - 1, 2**63 - 1, 2**63 - 1)]))           # every field at its upper edge
    pos.append((5, -2**63, [(0, -2**63, -2**63)]))                                 # ... and at its lower edge
    out, st = ctx.position_hash(*_csr(pos))
    assert not st.any()
    assert limbs_to_ints(out) == [state_tree.position_hash(pk, c, a) for pk, c, a in pos]
    # bounds: asset_id >= 2^120, public key >= p
    bad = [(7, 1, [(2**120, 1, 1)]), (P, 1, []), (3, 1, [(5, 1, 1)])]
    pk, col, off, aid, bal, fi = _csr(bad)
    out, st = ctx.position_hash(pk, col, off, aid, bal, fi)
    assert st.tolist() == [1, 1, 0] and limbs_to_ints(out)[:2] == [0, 0]


def test_multi_update_vs_dense_tree(ctx):
    """height-10 tree: the sparse update's previous root is the dense tree's root; its new root is the root of the dense tree
    rebuilt with the new leaves; inner nodes agree with the dense tree's."""
    height, n = 10, 37
    size = 1 << height
    leaves = rand_felts(size, 81)
    root, nodes, st = ctx.pedersen_merkle_tree(leaves, want_nodes=True)
    assert st == 0
    levels, off = [leaves], 0
    m = size // 2
    while m >= 1:
        levels.append(nodes[off:off + m]); off += m; m //= 2
    rng = random.Random(9)
    keys = sorted(rng.sample(range(size), n))
    new_vals = rand_felts(n, 82)
    sib = ctx.merkle_update_siblings(height, keys)
    sib_vals = np.array([levels[l][i] for l, i in sib], dtype=np.uint64).reshape(-1, 4)
    pr, nr, st, per_level = ctx.merkle_multi_update(height, keys, leaves[keys], new_vals, sib_vals, want_nodes=True)
    assert st == 0 and np.array_equal(pr, root)
    leaves2 = leaves.copy(); leaves2[keys] = new_vals
    root2, nodes2, _ = ctx.pedersen_merkle_tree(leaves2, want_nodes=True)
    assert np.array_equal(nr, root2)
    # inner nodes of the update tree, level by level: previous values = dense tree, new values = rebuilt dense tree
    off, m, cur = 0, size // 2, sorted(set(keys))
    for lvl in range(height):
        cur = sorted(set(k >> 1 for k in cur))
        assert per_level[lvl].shape == (2, len(cur), 4)
        assert np.array_equal(per_level[lvl][0], nodes[off:off + m][cur])
        assert np.array_equal(per_level[lvl][1], nodes2[off:off + m][cur])
        off += m; m //= 2
    # the sibling order is the oracle's
    upd = {k: (0, 0) for k in keys}
    _p, _n, used = state_tree.merkle_multi_update(height, upd, lambda l, i: 0, lambda a, b: 0)
    assert used == sib


def test_multi_update_height_64_vs_oracle(ctx):
    height, n = 64, 3
    rng = random.Random(64)
    keys = sorted(rng.randrange(2**64) for _ in range(n))
    keys[1] = keys[0] ^ 1 if keys[0] ^ 1 > keys[0] else keys[1]            # two leaves under one parent
    keys = sorted(set(keys))
    prev = [rng.randrange(P) for _ in keys]
    new = [rng.randrange(P) for _ in keys]
    sib = ctx.merkle_update_siblings(height, keys)
    table = {s: rng.randrange(P) for s in sib}
    pr, nr, st, _ = ctx.merkle_multi_update(height, keys, ints_to_limbs(prev), ints_to_limbs(new),
                                            ints_to_limbs([table[s] for s in sib]))
    want_p, want_n, used = state_tree.merkle_multi_update(height, {k: (p, v) for k, p, v in zip(keys, prev, new)},
                                                          lambda l, i: table[(l, i)])
    assert st == 0 and used == sib
    assert limbs_to_ints(pr.reshape(1, 4))[0] == want_p and limbs_to_ints(nr.reshape(1, 4))[0] == want_n


def test_multi_update_rejects_bad_keys_and_reports_status(ctx):
    from stark_perpetual_b200 import SpgError
    v = rand_felts(2, 5)
    with pytest.raises(SpgError, match="strictly increasing"):
        ctx.merkle_multi_update(8, [5, 5], v, v, rand_felts(14, 6))
    with pytest.raises(SpgError, match="outside the tree"):
        ctx.merkle_multi_update(8, [5, 256], v, v, rand_felts(14, 6))
    sib = ctx.merkle_update_siblings(8, [5, 9])
    with pytest.raises(SpgError, match="number of siblings"):
        ctx.merkle_multi_update(8, [5, 9], v, v, rand_felts(len(sib) + 1, 6))
    bad = v.copy(); bad[0] = ints_to_limbs([P])[0]
    _p, _n, st, _ = ctx.merkle_multi_update(8, [5, 9], bad, v, rand_felts(len(sib), 6))
    assert st == 1


def test_state_update_throughput_2_16(ctx, capsys):
    """2^16 position updates of a height-64 tree (the shape of state.cairo:143-173 for a large batch): hash every touched
    position, then the sparse multi-update; timed, and the new root cross-checked by updating in two halves."""
    rng = random.Random(16)
    n = 1 << 16
    g = np.random.Generator(np.random.PCG64(16))
    keys = np.unique(g.integers(0, 2**63, size=n + 1000, dtype=np.uint64))[:n]
    assert keys.shape[0] == n
    pos = _positions(rng, 2048, max_assets=4)
    pk, col, off, aid, bal, fi = _csr(pos)
    reps = n // 2048
    pkb, colb = np.tile(pk, (reps, 1)), np.tile(col, reps)
    counts = np.tile(np.diff(off.astype(np.int64)), reps)
    offb = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
    aidb, balb, fib = np.tile(aid, (reps, 1)), np.tile(bal, reps), np.tile(fi, reps)
    t0 = time.perf_counter()
    new_leaves, st = ctx.position_hash(pkb, colb, offb, aidb, balb, fib)
    t_hash, k_hash = time.perf_counter() - t0, ctx.last_kernel_ms
    assert not st.any()
    prev_leaves = rand_felts(n, 17)
    sib = ctx.merkle_update_siblings(64, keys)
    sib_vals = rand_felts(len(sib), 18)
    t0 = time.perf_counter()
    pr, nr, st, _ = ctx.merkle_multi_update(64, keys, prev_leaves, new_leaves, sib_vals)
    t_upd, k_upd = time.perf_counter() - t0, ctx.last_kernel_ms
    assert st == 0
    n_hashes = 2 * sum(len(set(int(k) >> l for k in keys[::1])) for l in (1,)) if False else None
    with capsys.disabled():
        print("\n[f-4] 2^16 updates, height 64: position_hash %.1f ms kernel / %.1f ms e2e; multi-update %.1f ms kernel / %.1f ms e2e, "
              "%d siblings" % (k_hash, 1e3 * t_hash, k_upd, 1e3 * t_upd, len(sib)))
    # consistency: prev_root of (new state) -> applying no change gives the same root on both planes
    pr2, nr2, st2, _ = ctx.merkle_multi_update(64, keys, new_leaves, new_leaves, sib_vals)
    assert st2 == 0 and np.array_equal(pr2, nr) and np.array_equal(nr2, nr)


def test_program_hash_chain_vs_oracle(ctx):
    """Row f-2: the right-folded chain behind compute_program_hash_chain (program_hash_test_utils.py:7-9) through the
    compat modules, against the oracle restatement, on a synthetic compiled program."""
    import os
    import sys
    from oracle import hash_chain as ohc
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "stark_perpetual_b200", "compat")
    sys.path.insert(0, compat)
    try:
        from starkware.cairo.bootloaders.hash_program import compute_program_hash_chain
        from starkware.cairo.common.hash_chain import compute_hash_chain
    finally:
        sys.path.remove(compat)
    rng = random.Random(2)
    data = [rng.randrange(P) for _ in range(40)]
    assert compute_hash_chain(data) == ohc.compute_hash_chain(data)
    assert compute_hash_chain(data[:1]) == data[0] and compute_hash_chain(data[:2]) == pedersen_hash(data[0], data[1])
    builtins = ["output", "pedersen", "range_check", "ecdsa", "bitwise"]      # main.cairo:1 of the perpetual program
    prog = {"builtins": builtins, "data": [hex(v) for v in data], "identifiers": {"__main__.main": {"pc": 17}}}
    assert compute_program_hash_chain(prog) == ohc.compute_program_hash_chain(builtins, 17, data)

    class Prog:
        pass
    p = Prog(); p.builtins, p.main, p.data = builtins, 17, data
    assert compute_program_hash_chain(p) == compute_program_hash_chain(prog)
    # batch of chains + range status
    arr = ints_to_limbs(data[:12] + data[12:23] + [P])
    out, st = ctx.hash_chain_rfold(arr, 12)
    assert st.tolist() == [0, 1] and limbs_to_ints(out)[0] == ohc.compute_hash_chain(data[:12])
