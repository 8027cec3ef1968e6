"""CPU: oracle/state_tree.py -- the update-tree shape against the reference's own build_update_tree
(src/starkware/python/merkle_tree.py:4-44, imported when reachable), the sparse multi-update against a full-tree
recomputation, position_hash against a direct restatement of hash.cairo:22-74 with a stand-in hash."""
import os
import random
import sys

import pytest

from oracle import refenv, state_tree


def cheap_hash(a, b):
    return (a * 1000003 + b * 7 + 12345) % (2**61 - 1)


def test_update_tree_shape_equals_reference():
    src = refenv.ref_src()
    if src is None:
        pytest.skip("reference not reachable")
    sys.path.insert(0, src)
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_merkle_tree", os.path.join(src, "starkware", "python", "merkle_tree.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        sys.path.remove(src)
    rng = random.Random(3)
    for height in (1, 2, 5, 16, 64):
        for n in (1, 2, 7, 40):
            mods = [(rng.randrange(1 << height), rng.randrange(1000)) for _ in range(n)]
            assert state_tree.build_update_tree(height, mods) == ref.build_update_tree(height, mods)
    assert state_tree.build_update_tree(5, []) is None and ref.build_update_tree(5, []) is None


@pytest.mark.parametrize("height,n", [(1, 1), (3, 2), (6, 5), (8, 30), (8, 256)])
def test_multi_update_equals_full_tree_recomputation(height, n):
    rng = random.Random(height * 100 + n)
    size = 1 << height
    leaves = [rng.randrange(10**9) for _ in range(size)]
    levels = [leaves]
    while len(levels[-1]) > 1:
        p = levels[-1]
        levels.append([cheap_hash(p[2 * i], p[2 * i + 1]) for i in range(len(p) // 2)])
    keys = sorted(rng.sample(range(size), n))
    updates = {k: (leaves[k], rng.randrange(10**9)) for k in keys}
    prev_root, new_root, used = state_tree.merkle_multi_update(height, updates, lambda l, i: levels[l][i], cheap_hash)
    assert prev_root == levels[-1][0]
    new_leaves = list(leaves)
    for k, (_p, v) in updates.items():
        new_leaves[k] = v
    lv = new_leaves
    while len(lv) > 1:
        lv = [cheap_hash(lv[2 * i], lv[2 * i + 1]) for i in range(len(lv) // 2)]
    assert new_root == lv[0]
    # every sibling hangs off the update tree: it is never an ancestor of an updated leaf
    for level, index in used:
        assert all((k >> level) != index for k in keys)


def test_position_hash_structure():
    calls = []

    def h(a, b):
        calls.append((a, b))
        return cheap_hash(a, b)
    assets = [(5, -3, 7), (2**120 - 1, 2**63 - 1, -(2**63))]
    out = state_tree.position_hash(1234, -1, assets, h)
    assert len(calls) == len(assets) + 2 and calls[0][0] == 0
    assert calls[0][1] == (5 * 2**64 + (7 + 2**63)) * 2**64 + (-3 + 2**63)
    assert calls[1][1] == ((2**120 - 1) * 2**64 + 0) * 2**64 + (2**64 - 1)
    assert calls[2][1] == 1234 and calls[3][1] == (2**63 - 1) * 2**16 + 2 and out == cheap_hash(*calls[3])
    assert calls[1][1] < 2**251
    # empty position: H(H(0, key), (collateral + 2^63) * 2^16)
    assert state_tree.position_hash(9, 0, [], cheap_hash) == cheap_hash(cheap_hash(0, 9), 2**63 * 2**16)


def test_hash_chain_is_a_right_fold():
    from oracle import hash_chain
    calls = []

    def h(a, b):
        calls.append((a, b))
        return cheap_hash(a, b)
    data = [11, 22, 33, 44]
    out = hash_chain.compute_hash_chain(data, h)
    assert calls == [(33, 44), (22, cheap_hash(33, 44)), (11, cheap_hash(22, cheap_hash(33, 44)))] and out == cheap_hash(*calls[-1])
    assert hash_chain.compute_hash_chain([7], h) == 7
    # program header: [len, bootloader_version, main, n_builtins, builtins..., data...]
    calls.clear()
    hash_chain.compute_program_hash_chain(["output", "pedersen"], 5, [100, 200], hash_func=h)
    first_args = [c[0] for c in calls][::-1]
    assert first_args == [7, 0, 5, 2, int.from_bytes(b"output", "big"), int.from_bytes(b"pedersen", "big"), 100]
