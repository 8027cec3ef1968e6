"""GPU parity: batched ECDSA verification / key derivation / the reference-compatible signature module against
golden vectors generated from the reference (signature.py:217-260) and the Python oracle -- exact statuses."""
import os
import random
import sys

import numpy as np
import pytest

from oracle import ecdsa as oecdsa
from oracle.params import EC_ORDER, FIELD_PRIME as P
from stark_perpetual_b200._lib import ints_to_limbs, limbs_to_ints

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def compat():
    sys.path.insert(0, os.path.join(ROOT, "stark_perpetual_b200", "compat"))
    from starkware.crypto.signature import signature
    return signature


def test_verify_golden_all_statuses(ctx, golden):
    xonly = [v for v in golden["verify"] if isinstance(v[3], str)]
    point = [v for v in golden["verify"] if not isinstance(v[3], str)]
    st = ctx.ecdsa_verify(ints_to_limbs([int(v[0], 16) for v in xonly]), ints_to_limbs([int(v[1], 16) for v in xonly]),
                          ints_to_limbs([int(v[2], 16) for v in xonly]), ints_to_limbs([int(v[3], 16) for v in xonly]))
    assert list(st) == [v[4] for v in xonly], [v[5] for v in xonly]
    st = ctx.ecdsa_verify(ints_to_limbs([int(v[0], 16) for v in point]), ints_to_limbs([int(v[1], 16) for v in point]),
                          ints_to_limbs([int(v[2], 16) for v in point]), ints_to_limbs([int(v[3][0], 16) for v in point]),
                          ints_to_limbs([int(v[3][1], 16) for v in point]))
    assert list(st) == [v[4] for v in point], [v[5] for v in point]


def test_private_to_public_golden(ctx, golden):
    privs = [int(a, 16) for a, _b in golden["keys"]]
    out, st = ctx.private_to_stark_key(ints_to_limbs(privs))
    assert not st.any()
    assert limbs_to_ints(out) == [int(b, 16) for _a, b in golden["keys"]]
    # large keys (bit 251 set) and range statuses
    ks = [EC_ORDER - 1, EC_ORDER - 2, 2**251, 1, 2]
    x, y, st = ctx.private_to_stark_key(ints_to_limbs(ks), want_y=True)
    assert not st.any()
    for k, gx, gy in zip(ks, limbs_to_ints(x), limbs_to_ints(y)):
        assert (gx, gy) == oecdsa.private_key_to_ec_point_on_stark_curve(k)
    _o, st = ctx.private_to_stark_key(ints_to_limbs([0, EC_ORDER, 2**256 - 1]))
    assert list(st) == [1, 1, 1]


def make_orders(n, n_keys, seed, corrupt_every=0):
    rng = random.Random(seed)
    privs = [rng.randrange(1, EC_ORDER) for _ in range(n_keys)]
    pubs = [oecdsa.private_to_stark_key(k) for k in privs]
    msgs, rs, ss, keys, want = [], [], [], [], []
    for i in range(n):
        k = i % n_keys
        m = rng.randrange(2**251)
        r, s = oecdsa.sign(m, privs[k])
        ok = True
        if corrupt_every and i % corrupt_every == 0:
            which = rng.randrange(4)
            if which == 0:
                m ^= 1 << rng.randrange(250)
            elif which == 1:
                r ^= 1 << rng.randrange(250)
                r = r or 1
            elif which == 2:
                s = (s ^ (1 << rng.randrange(250))) % EC_ORDER or 1
            else:
                k = (k + 1) % n_keys
            ok = False
        msgs.append(m); rs.append(r); ss.append(s); keys.append(pubs[k]); want.append(ok)
    return msgs, rs, ss, keys, want


def test_batch_verify_synthetic_orders(ctx):
    """Config-5 shape at a size the Python oracle signs in seconds: 96 signatures, 8 keys, every 5th corrupted;
    statuses equal the oracle's verify on every element."""
    msgs, rs, ss, keys, want = make_orders(96, 8, 1005, corrupt_every=5)
    st = ctx.ecdsa_verify(ints_to_limbs(msgs), ints_to_limbs(rs), ints_to_limbs(ss), ints_to_limbs(keys))
    got = [int(v) for v in st]
    for i in range(96):
        try:
            ref = 1 if oecdsa.verify(msgs[i], rs[i], ss[i], keys[i]) else 0
        except AssertionError:
            ref = 2
        assert got[i] == ref, i
    assert sum(got) >= 70


def test_batch_verify_large_replicated(ctx):
    """65536 signatures (BASELINE.json configs[4] size) built by replicating 64 distinct ones: every copy must
    get the status of its original (determinism across the whole grid)."""
    msgs, rs, ss, keys, want = make_orders(64, 4, 77, corrupt_every=7)
    reps = 1024
    st = ctx.ecdsa_verify(np.tile(ints_to_limbs(msgs), (reps, 1)), np.tile(ints_to_limbs(rs), (reps, 1)),
                          np.tile(ints_to_limbs(ss), (reps, 1)), np.tile(ints_to_limbs(keys), (reps, 1)))
    base = [1 if w else 0 for w in want]
    assert st.reshape(reps, 64).tolist() == [base] * reps


def test_compat_signature_module(ctx, golden):
    sig = compat()
    # Pedersen KATs of the reference (signature_test_data.json:190-201 via the golden file)
    for a, b, o, tag in golden["pedersen"][:6]:
        assert sig.pedersen_hash(int(a, 16), int(b, 16)) == int(o, 16)
    with pytest.raises(AssertionError):
        sig.pedersen_hash(P, 1)
    # deterministic signing: the 4 JS RFC 6979 KATs (signature.spec.js:96-137)
    for mh, priv, er, es in golden["sign_js_kat"]:
        assert sig.sign(int(mh, 16), int(priv, 16)) == (int(er, 16), int(es, 16))
    for msg, priv, r, s in golden["sign"][:4]:
        assert sig.sign(int(msg, 16), int(priv, 16)) == (int(r, 16), int(s, 16))
        pub = sig.private_to_stark_key(int(priv, 16))
        assert sig.verify(int(msg, 16), int(r, 16), int(s, 16), pub)
        assert sig.verify(int(msg, 16), int(r, 16), int(s, 16), sig.private_key_to_ec_point_on_stark_curve(int(priv, 16)))
        assert not sig.verify(int(msg, 16) ^ 1, int(r, 16), int(s, 16), pub)
    # exceptions exactly where the reference raises
    for msg, r, s, pub, res, tag in golden["verify"]:
        pk = int(pub, 16) if isinstance(pub, str) else (int(pub[0], 16), int(pub[1], 16))
        if res == 2:
            with pytest.raises(AssertionError):
                sig.verify(int(msg, 16), int(r, 16), int(s, 16), pk)
        else:
            assert sig.verify(int(msg, 16), int(r, 16), int(s, 16), pk) == bool(res), tag


def test_sign_batch_matches_reference_vectors_and_scalar_sign(ctx, golden):
    """sign_batch (spg_sign_batch: nonce, k*G and the mod-n finish all in the kernel) against the 4 JS KATs
    (signature.spec.js:96-137), the reference-generated sign vectors and the oracle's sign."""
    sig = compat()
    kat = golden["sign_js_kat"]
    vec = golden["sign"]
    msgs = [int(v[0], 16) for v in kat] + [int(v[0], 16) for v in vec]
    privs = [int(v[1], 16) for v in kat] + [int(v[1], 16) for v in vec]
    want = [(int(v[2], 16), int(v[3], 16)) for v in kat] + [(int(v[2], 16), int(v[3], 16)) for v in vec]
    assert sig.sign_batch(msgs, privs) == want
    rng = random.Random(31)
    m2 = [rng.randrange(2**251) for _ in range(64)]
    k2 = [rng.randrange(1, EC_ORDER) for _ in range(64)]
    got = sig.sign_batch(m2, k2)
    assert got[:8] == [oecdsa.sign(m, k) for m, k in zip(m2[:8], k2[:8])]
    pubs = sig.private_to_stark_key_batch(k2)
    assert sig.verify_batch(m2, [g[0] for g in got], [g[1] for g in got], pubs) == [True] * 64
    assert sig.sign_batch(m2[:3], k2[:3], seeds=[5, None, 7])[0] == oecdsa.sign(m2[0], k2[0], 5)
    # scalar sign: in-domain arguments go through the kernel, a 70-bit seed through the host loop; both the reference's
    for seed in (None, 0, 1, 2**40 + 3, 2**70 + 1):
        assert sig.sign(m2[1], k2[1], seed) == oecdsa.sign(m2[1], k2[1], seed)
    # messages around the nibble rule (signature.py:119-121) and the C-ABI status codes
    edge = [0, 1, 2**248 - 1, 2**248, 2**249 + 3, 2**251 - 1]
    assert sig.sign_batch(edge, k2[:6]) == [oecdsa.sign(m, k) for m, k in zip(edge, k2[:6])]
    _r, _s, st = ctx.sign(ints_to_limbs([2**251, 5, 5, 5]), ints_to_limbs([5, 0, EC_ORDER, 7]))
    assert st.tolist() == [1, 2, 2, 0]
    with pytest.raises(AssertionError):
        sig.sign(2**251, 5)


def test_pedersen_merkle_tree(ctx):
    """Merkle tree with pedersen_hash nodes (StarkEx state-tree node function) against the oracle's hash, node by node."""
    from oracle.pedersen import pedersen_hash as oph
    sig = compat()
    rng = random.Random(99)
    leaves = [rng.randrange(P) for _ in range(32)]
    leaves[0], leaves[1], leaves[2] = 0, P - 1, 1
    root, nodes, st = ctx.pedersen_merkle_tree(ints_to_limbs(leaves), want_nodes=True)
    assert st == 0
    level, want_nodes = leaves, []
    while len(level) > 1:
        level = [oph(level[2 * i], level[2 * i + 1]) for i in range(len(level) // 2)]
        want_nodes += level
    assert limbs_to_ints(nodes) == want_nodes
    assert limbs_to_ints(root.reshape(1, 4))[0] == want_nodes[-1] == sig.pedersen_merkle_root(leaves)
    # two leaves: the tree is one hash; a leaf >= p is reported
    r2, _n, st = ctx.pedersen_merkle_tree(ints_to_limbs([3, 4]))
    assert st == 0 and limbs_to_ints(r2.reshape(1, 4))[0] == oph(3, 4)
    _r, _n, st = ctx.pedersen_merkle_tree(ints_to_limbs([3, P, 5, 6]))
    assert st == 1


def test_remaining_signature_module_functions(ctx, golden):
    """get_y_coordinate / is_valid_stark_key (reference-generated vectors), pedersen_hash_as_point, mimic_ec_mult_air and
    the fast_pedersen_hash byte ABI, through the compat module, against the oracle."""
    from oracle import pedersen as opedersen
    from oracle.params import EC_GEN, SHIFT_POINT
    sig = compat()
    from starkware.crypto.signature import fast_pedersen_hash as fph
    for x, y in golden["get_y"]:
        if y is None:
            with pytest.raises(sig.InvalidPublicKeyError):
                sig.get_y_coordinate(int(x, 16))
            assert not sig.is_valid_stark_key(int(x, 16))
        else:
            assert sig.get_y_coordinate(int(x, 16)) == int(y, 16)
            assert sig.is_valid_stark_key(int(x, 16))
    xs = [int(x, 16) for x, _y in golden["get_y"]] + [0, 1, P - 1]
    ys, st = ctx.get_y_coordinate(ints_to_limbs(xs))
    for x, y, s in zip(xs, limbs_to_ints(ys), st):
        try:
            assert (int(s), y) == (0, oecdsa.get_y_coordinate(x))
        except oecdsa.InvalidPublicKeyError:
            assert int(s) == 1
    assert ctx.get_y_coordinate(ints_to_limbs([P]))[1][0] == 2
    # hash as point
    for a, b, _o, _t in golden["pedersen"][:6]:
        assert sig.pedersen_hash_as_point(int(a, 16), int(b, 16)) == opedersen.pedersen_hash_as_point(int(a, 16), int(b, 16))
    assert sig.pedersen_hash_as_point(5) == opedersen.pedersen_hash_as_point(5)
    # mimic_ec_mult_air: values and the assertion cases
    rng = random.Random(5)
    q = oecdsa.private_key_to_ec_point_on_stark_curve(rng.randrange(1, EC_ORDER))
    for m in (1, 2, 3, rng.randrange(1, 2**251), 2**251 - 1):
        assert sig.mimic_ec_mult_air(m, q, SHIFT_POINT) == oecdsa.mimic_ec_mult_air(m, q, SHIFT_POINT)
    assert sig.mimic_ec_mult_air(7, EC_GEN, q) == oecdsa.mimic_ec_mult_air(7, EC_GEN, q)
    for bad in ((0, q, SHIFT_POINT), (2**251, q, SHIFT_POINT), (5, q, q), (5, q, (q[0], P - q[1]))):
        with pytest.raises(AssertionError):
            sig.mimic_ec_mult_air(*bad)
        with pytest.raises(AssertionError):
            oecdsa.mimic_ec_mult_air(*bad)
    assert sig.is_point_on_curve(*q) and not sig.is_point_on_curve(q[0], q[1] + 1)
    assert sig.is_valid_stark_private_key(1) and not sig.is_valid_stark_private_key(EC_ORDER)
    assert sig.is_valid_stark_private_key(sig.get_random_private_key())
    # byte ABI of fast_pedersen_hash
    a, b = int(golden["pedersen"][0][0], 16), int(golden["pedersen"][0][1], 16)
    assert fph.pedersen_hash(a, b) == int(golden["pedersen"][0][2], 16)
    assert fph.pedersen_hash_func(a.to_bytes(32, "big"), b.to_bytes(32, "big")) == int(golden["pedersen"][0][2], 16).to_bytes(32, "big")
    assert fph.pedersen_hash_func_batch(a.to_bytes(32, "big") * 3, b.to_bytes(32, "big") * 3) == int(golden["pedersen"][0][2], 16).to_bytes(32, "big") * 3
