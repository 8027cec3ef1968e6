"""State-tree work of a perpetual batch, plain Python.  TEST INFRASTRUCTURE.

Restates
  * position_hash_assets / position_hash      src/services/perpetual/cairo/position/hash.cairo:22-74
    (constants: src/services/perpetual/cairo/definitions/constants.cairo:10-38),
  * the sparse multi-update that merkle_multi_update performs for the positions / orders trees
    (src/services/perpetual/cairo/state/state.cairo:143-173), over the update tree of
    src/starkware/python/merkle_tree.py:4-44 (build_update_tree / decode_node).
Parity: the hash function is pinned (oracle/pedersen.py vs the reference's vectors); the tree shape is checked against the
reference's own build_update_tree in tests/test_oracle_state_tree.py whenever the reference is reachable; the reference holds
no known-answer vector for position_hash itself (the Cairo code is its only statement), so that packing is "restated, no KAT".
"""
from .pedersen import pedersen_hash

BALANCE_LOWER_BOUND = -(2**63)
BALANCE_UPPER_BOUND = 2**63
FUNDING_INDEX_LOWER_BOUND = -(2**63)
FUNDING_INDEX_UPPER_BOUND = 2**63
N_ASSETS_UPPER_BOUND = 2**16
ASSET_ID_UPPER_BOUND = 2**120


def position_hash(public_key, collateral_balance, assets, hash_func=pedersen_hash):
    """assets: iterable of (asset_id, balance, cached_funding_index), sorted by asset_id (hash.cairo:57-59)."""
    assets = list(assets)
    h = 0
    for asset_id, balance, funding_index in assets:                              # hash.cairo:22-45
        assert 0 <= asset_id < ASSET_ID_UPPER_BOUND
        assert FUNDING_INDEX_LOWER_BOUND <= funding_index < FUNDING_INDEX_UPPER_BOUND
        assert BALANCE_LOWER_BOUND <= balance < BALANCE_UPPER_BOUND
        packed = asset_id
        packed = packed * (FUNDING_INDEX_UPPER_BOUND - FUNDING_INDEX_LOWER_BOUND) + (funding_index - FUNDING_INDEX_LOWER_BOUND)
        packed = packed * (BALANCE_UPPER_BOUND - BALANCE_LOWER_BOUND) + (balance - BALANCE_LOWER_BOUND)
        h = hash_func(h, packed)
    assert len(assets) < N_ASSETS_UPPER_BOUND
    h = hash_func(h, public_key)                                                  # hash.cairo:66
    return hash_func(h, (collateral_balance - BALANCE_LOWER_BOUND) * N_ASSETS_UPPER_BOUND + len(assets))   # :69-73


def build_update_tree(height, modifications):
    """The subtree induced by the modified leaves (merkle_tree.py:4-29): None, a (left, right) pair, or a leaf value."""
    if len(modifications) == 0:
        return None
    layer = dict(modifications)
    for _ in range(height):
        parents = set(index // 2 for index in layer)
        layer = {index: (layer.get(index * 2), layer.get(index * 2 + 1)) for index in parents}
    assert len(layer) == 1
    return layer[0]


def merkle_multi_update(height, updates, sibling, hash_func=pedersen_hash):
    """updates: {leaf index: (prev_value, new_value)}; sibling(level, index) -> hash of the untouched subtree rooted at node
    `index` of level `level` (0 = leaves).  Returns (prev_root, new_root, [(level, index)] of the siblings used, in the
    order of include/spg.h: bottom-up by level, ascending index)."""
    tree = build_update_tree(height, [(k, (k, v)) for k, v in updates.items()])
    used = []

    def walk(node, level, index):
        if level == 0:
            _k, (prev, new) = node
            assert _k == index
            return prev, new
        left, right = node                                                        # decode_node (merkle_tree.py:32-44)
        assert left is not None or right is not None, "No updates in tree"
        if left is None:
            s = sibling(level - 1, 2 * index)
            used.append((level - 1, 2 * index))
            lp = ln = s
        else:
            lp, ln = walk(left, level - 1, 2 * index)
        if right is None:
            s = sibling(level - 1, 2 * index + 1)
            used.append((level - 1, 2 * index + 1))
            rp = rn = s
        else:
            rp, rn = walk(right, level - 1, 2 * index + 1)
        return hash_func(lp, rp), hash_func(ln, rn)
    prev_root, new_root = walk(tree, height, 0)
    return prev_root, new_root, sorted(used)
