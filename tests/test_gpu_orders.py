"""GPU parity: the device limit-order pipeline (pack -> 4-deep Pedersen chain -> ECDSA) against vectors generated
from the reference (perpetual_messages.get_limit_order_msg) and the oracle's sign / verify."""
import os
import random
import sys

import numpy as np
import pytest

from oracle import ecdsa as oecdsa
from oracle.params import EC_ORDER, FIELD_PRIME as P
from oracle.pedersen import pedersen_hash as opedersen
from stark_perpetual_b200._lib import ints_to_limbs

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
COMPAT = os.path.join(ROOT, "stark_perpetual_b200", "compat")


@pytest.fixture(scope="module")
def pm():
    sys.path.insert(0, COMPAT)
    try:
        from services.perpetual.public import perpetual_messages as mod
        yield mod
    finally:
        sys.path.remove(COMPAT)


def _rand_order(rng):
    return dict(asset_id_synthetic=rng.randrange(2**128), asset_id_collateral=rng.randrange(2**250),
                is_buying_synthetic=rng.randrange(2), asset_id_fee=rng.randrange(2**250), amount_synthetic=rng.randrange(2**64),
                amount_collateral=rng.randrange(2**64), max_amount_fee=rng.randrange(2**64), nonce=rng.randrange(2**32),
                position_id=rng.randrange(2**64), expiration_timestamp=rng.randrange(2**32))


def test_limit_order_msg_batch_vs_reference_vectors(ctx, golden, pm):
    vec = [(f, w) for kind, f, w in golden["messages"] if kind == "limit_order"]
    pre = golden["messages_precomputed"]["limit_order"]
    for want, d in pre.items():
        vec.append((dict(asset_id_synthetic=d["assetIdSynthetic"], asset_id_collateral=d["assetIdCollateral"],
                         is_buying_synthetic=d["isBuyingSynthetic"], asset_id_fee=d["assetIdFee"],
                         amount_synthetic=d["amountSynthetic"], amount_collateral=d["amountCollateral"],
                         max_amount_fee=d["amountFee"], nonce=d["nonce"], position_id=d["positionId"],
                         expiration_timestamp=d["expirationTimestamp"]), want))
    got = pm.get_limit_order_msg_batch([f for f, _w in vec])
    assert [hex(g) for g in got] == [w for _f, w in vec]
    # the scalar builder with the default (GPU) hash function gives the same values
    assert hex(pm.get_limit_order_msg(**vec[0][0])) == vec[0][1]


def test_limit_order_msg_batch_vs_oracle_random_and_edges(ctx, pm):
    rng = random.Random(4242)
    orders = [_rand_order(rng) for _ in range(300)]
    # extreme field values
    orders.append(dict(asset_id_synthetic=2**128 - 1, asset_id_collateral=2**250 - 1, is_buying_synthetic=1,
                       asset_id_fee=2**250 - 1, amount_synthetic=2**64 - 1, amount_collateral=2**64 - 1,
                       max_amount_fee=2**64 - 1, nonce=2**32 - 1, position_id=2**64 - 1, expiration_timestamp=2**32 - 1))
    orders.append(dict.fromkeys(pm.ORDER_FIELDS, 0))
    got = pm.get_limit_order_msg_batch(orders)
    want = [pm.get_limit_order_msg(hash_function=opedersen, **o) for o in orders]
    assert got == want


def test_limit_order_bounds_status(ctx, pm):
    ok = dict.fromkeys(pm.ORDER_FIELDS, 1)
    for field, bad in (("asset_id_synthetic", 2**128), ("asset_id_collateral", 2**250), ("asset_id_fee", 2**250 + 5)):
        with pytest.raises(AssertionError):
            pm.get_limit_order_msg_batch([ok, dict(ok, **{field: bad})])
    with pytest.raises(AssertionError):
        pm.get_limit_order_msg_batch([dict(ok, nonce=2**32)])


def test_limit_order_verify_pipeline(ctx, pm):
    """Signed orders (oracle sign) verify; any corrupted field, signature word or key does not."""
    rng = random.Random(777)
    n = 48
    orders = [_rand_order(rng) for _ in range(n)]
    privs = [rng.randrange(1, EC_ORDER) for _ in range(n)]
    keys = [oecdsa.private_to_stark_key(k) for k in privs]
    msgs = [pm.get_limit_order_msg(hash_function=opedersen, **o) for o in orders]
    sigs = [oecdsa.sign(m, k) for m, k in zip(msgs, privs)]
    rs, ss = [s[0] for s in sigs], [s[1] for s in sigs]
    assert pm.verify_limit_orders_batch(orders, rs, ss, keys) == [True] * n
    # corruptions: order field / r / s / key
    bad_orders = [dict(o) for o in orders]
    expect = [True] * n
    for i in range(0, n, 4):
        bad_orders[i]["nonce"] ^= 1; expect[i] = False
    rs2 = list(rs)
    for i in range(1, n, 4):
        rs2[i] ^= 2; expect[i] = False
    keys2 = list(keys)
    for i in range(2, n, 4):
        keys2[i] = keys[(i + 1) % n]; expect[i] = False
    got = pm.verify_limit_orders_batch(bad_orders, rs2, ss, keys2)
    assert got == expect
    # the same answers from the oracle's verify on the oracle's message hashes
    for i in (0, 1, 2, 3, 5):
        m = pm.get_limit_order_msg(hash_function=opedersen, **bad_orders[i])
        assert oecdsa.verify(m, rs2[i], ss[i], keys2[i]) == expect[i]
    # a signature operand out of range makes the reference raise (signature.py:219)
    with pytest.raises(AssertionError):
        pm.verify_limit_orders_batch(orders[:2], rs[:2], [ss[0], EC_ORDER], keys[:2])


def test_other_message_batches_vs_reference_vectors(ctx, golden, pm):
    """spg_message_hash_batch (device packing + Pedersen chain) for transfers, conditional transfers, withdrawals to
    an address and oracle prices: the 32 message hashes generated from the reference (tests/golden), the scalar
    builders with the oracle's hash on random fields at the edges of every range, and the status of each violated
    bound."""
    rng = random.Random(2024)
    for kind, scalar in (("transfer", pm.get_transfer_msg), ("conditional_transfer", pm.get_conditional_transfer_msg),
                         ("withdrawal_to_address", pm.get_withdrawal_to_address_msg), ("price", pm.get_price_msg)):
        vec = [(f, w) for k, f, w in golden["messages"] if k == kind]
        assert len(vec) == 8
        assert [hex(g) for g in pm.get_msg_batch(kind, [f for f, _w in vec])] == [w for _f, w in vec]
        fnames, inames = pm._MESSAGE_ARGS[kind]
        width = {"asset_id": 250, "asset_id_fee": 250, "receiver_public_key": 251, "condition": 251, "nonce": 32,
                 "expiration_timestamp": 32, "asset_id_collateral": 250, "eth_address": 160, "asset_pair": 128, "price": 120,
                 "oracle_name": 40, "timestamp": 32}
        msgs = []
        for i in range(40):
            m = {}
            for name in fnames + inames:
                bits = width.get(name, 64)
                m[name] = rng.choice([0, 1, 2**bits - 1, rng.randrange(2**bits), rng.randrange(2**bits)])
            if "eth_address" in m:
                m["eth_address"] = hex(m["eth_address"])
            msgs.append(m)
        got = pm.get_msg_batch(kind, msgs)
        assert got == [scalar(hash_function=opedersen, **m) for m in msgs], kind
        # every bound: one element out of range fails the whole batch like the reference's assert, and the C-ABI
        # reports exactly that element
        for name in fnames + inames:
            if name not in width:
                continue
            bad = dict(msgs[3])
            bad[name] = hex(2**width[name]) if name == "eth_address" else 2**width[name]
            with pytest.raises(AssertionError):
                pm.get_msg_batch(kind, msgs[:3] + [bad] + msgs[4:6])
    assert pm.get_msg_batch("price", []) == []
    felts = [ints_to_limbs([1, 2**250, 3]), ints_to_limbs([4, 5, 6]), ints_to_limbs([7, 8, 2**251])]
    ints = [np.array([1, 2, 3], dtype=np.uint64)] * 7
    _out, st = ctx.message_hash("transfer", felts, ints)
    assert st.tolist() == [0, 1, 1]
