"""CPU: the table-driven message builders of the compat mirror (stark_perpetual_b200/compat/services/...) against
the reference's own KATs (perpetual_messages_precomputed.json, copied into the golden file) and 40 vectors generated
from the reference, with the ORACLE's Pedersen hash injected through `hash_function=` (the reference's own
injection point, perpetual_messages.py:36...) -- so the packing / chaining logic is pinned without a GPU."""
import os
import sys

import pytest

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


def _oracle_hash():
    from oracle.pedersen import pedersen_hash
    return pedersen_hash


def test_message_builders_vs_reference_vectors(golden, pm):
    h = _oracle_hash()
    fn = {"limit_order": pm.get_limit_order_msg, "transfer": pm.get_transfer_msg,
          "conditional_transfer": pm.get_conditional_transfer_msg,
          "withdrawal_to_address": pm.get_withdrawal_to_address_msg, "price": pm.get_price_msg}
    for kind, fields, want in golden["messages"]:
        assert hex(fn[kind](hash_function=h, **fields)) == want, kind


def test_reference_kats(golden, pm):
    """The vectors of src/services/perpetual/public/perpetual_messages_test.py:22-88."""
    h, pre = _oracle_hash(), golden["messages_precomputed"]
    for want, d in pre["limit_order"].items():
        assert hex(pm.get_limit_order_msg(d["assetIdSynthetic"], d["assetIdCollateral"], d["isBuyingSynthetic"], d["assetIdFee"],
                                          d["amountSynthetic"], d["amountCollateral"], d["amountFee"], d["nonce"],
                                          d["positionId"], d["expirationTimestamp"], hash_function=h)) == want
    for want, d in pre["conditional_transfer"].items():
        assert hex(pm.get_conditional_transfer_msg(d["assetId"], d["assetIdFee"], d["receiverPublicKey"], d["condition"],
                                                   d["senderPositionId"], d["receiverPositionId"], d["srcFeePositionId"],
                                                   d["nonce"], d["amount"], d["maxAmountFee"], d["expirationTimestamp"],
                                                   hash_function=h)) == want
    for want, d in pre["transfer"].items():
        assert hex(pm.get_transfer_msg(d["assetId"], d["assetIdFee"], d["receiverPublicKey"], d["senderPositionId"],
                                       d["receiverPositionId"], d["feePositionId"], d["nonce"], d["amount"], d["maxAmountFee"],
                                       d["expirationTimestamp"], hash_function=h)) == want
    for want, d in pre["withdrawal_to_address"].items():
        assert hex(pm.get_withdrawal_to_address_msg(asset_id_collateral=d["assetIdCollateral"], eth_address=d["ethAddress"],
                                                    position_id=d["positionId"], nonce=d["nonce"],
                                                    expiration_timestamp=d["expirationTimestamp"], amount=d["amount"],
                                                    hash_function=h)) == want


def test_bounds_raise_like_the_reference(pm):
    h = _oracle_hash()
    ok = dict(asset_id_synthetic=1, asset_id_collateral=1, is_buying_synthetic=1, asset_id_fee=1, amount_synthetic=1,
              amount_collateral=1, max_amount_fee=1, nonce=1, position_id=1, expiration_timestamp=1)
    for field, bad in (("asset_id_synthetic", 2**128), ("asset_id_collateral", 2**250), ("asset_id_fee", 2**250),
                       ("amount_synthetic", 2**64), ("nonce", 2**32), ("position_id", -1), ("expiration_timestamp", 2**32)):
        with pytest.raises(AssertionError):
            pm.get_limit_order_msg(hash_function=h, **dict(ok, **{field: bad}))
    with pytest.raises(AssertionError):
        pm.get_price_msg(2**40, 1, 1, 1, hash_function=h)


def test_limit_order_elements_layout(pm):
    """The two packed words of a limit order, bit by bit (what k_pack_limit_orders writes)."""
    e = pm.limit_order_elements(asset_id_synthetic=7, asset_id_collateral=9, is_buying_synthetic=0, asset_id_fee=11,
                                amount_synthetic=0xAAAA, amount_collateral=0xBBBB, max_amount_fee=0xCCCC, nonce=0xDDDD,
                                position_id=0x1234, expiration_timestamp=0x5678)
    assert e[:3] == [7, 9, 11]
    assert e[3] == (0xAAAA << 160) | (0xBBBB << 96) | (0xCCCC << 32) | 0xDDDD
    assert e[4] == (3 << 241) | (0x1234 << 177) | (0x1234 << 113) | (0x1234 << 49) | (0x5678 << 17)
    e = pm.limit_order_elements(7, 9, 1, 11, 0xAAAA, 0xBBBB, 0xCCCC, 0xDDDD, 0x1234, 0x5678)
    assert e[:2] == [9, 7] and e[3] >> 160 == 0xBBBB


def test_compat_signature_host_side_pieces(golden):
    """The host-side parts of the compat signature module (no GPU needed): constants, grind_key (key_derivation.spec.js
    KAT + reference-generated vectors), the RFC 6979 nonce derivation against the oracle's, and the pure predicates."""
    sys.path.insert(0, COMPAT)
    try:
        from starkware.crypto.signature import signature as sig
    finally:
        sys.path.remove(COMPAT)
    from oracle import ecdsa as oecdsa, params
    assert (sig.FIELD_PRIME, sig.EC_ORDER, sig.ALPHA, sig.BETA, sig.FIELD_GEN) == (
        params.FIELD_PRIME, params.EC_ORDER, params.ALPHA, params.BETA, 3)
    assert sig.SHIFT_POINT == params.SHIFT_POINT and sig.EC_GEN == params.EC_GEN
    assert sig.N_ELEMENT_BITS_ECDSA == 251 and sig.N_ELEMENT_BITS_HASH == 252
    for a, b, o in golden["grind_key"]:
        assert sig.grind_key(int(a, 16), int(b, 16)) == int(o, 16)
    for mh, priv, _r, _s in golden["sign_js_kat"]:
        for seed in (None, 1, 77):
            assert sig.generate_k_rfc6979(int(mh, 16), int(priv, 16), seed) == oecdsa.generate_k_rfc6979(int(mh, 16), int(priv, 16), seed)
    assert sig.is_valid_stark_private_key(1) and not sig.is_valid_stark_private_key(0)
    assert sig.is_point_on_curve(*params.EC_GEN) and not sig.is_point_on_curve(1, 1)
    assert sig.inv_mod_curve_size(7) * 7 % params.EC_ORDER == 1
