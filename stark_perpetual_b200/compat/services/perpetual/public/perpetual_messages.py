"""GPU-backed mirror of the reference's `services.perpetual.public.perpetual_messages`
(src/services/perpetual/public/perpetual_messages.py): the same builder names, argument order, bounds
(AssertionError) and `hash_function=` injection point; the default hash is libspg's Pedersen kernel.

The builders are table-driven: every message is a chain  h = H(a, b); h = H(h, c); ...  over a few full field
elements followed by one or two PACKED words, and a packed word is a list of (value, bit width) fields, most
significant first, plus zero padding (reference :80-94, :138-162, :201-209, :266-286, :321-326; Cairo twin
src/services/exchange/cairo/signature_message_hashes.cairo:56-170).

`*_batch` functions are the device pipeline (spg_limit_order_msg_batch / spg_limit_order_verify_batch): packing,
the 4-deep hash chain and the signature check run on the GPU without returning to the host in between.
"""
import numpy as np

from starkware.crypto.signature.signature import pedersen_hash

LIMIT_ORDER_WITH_FEES = 3
TRANSFER = 4
CONDITIONAL_TRANSFER = 5
WITHDRAWAL_TO_ADDRESS = 7


def _pack(fields, pad_bits=0):
    """fields: [(value, bits)] most significant first."""
    word = 0
    for value, bits in fields:
        word = (word << bits) + value
    return word << pad_bits


def _chain(hash_function, elements):
    msg = hash_function(elements[0], elements[1])
    for e in elements[2:]:
        msg = hash_function(msg, e)
    return msg


def _check(bounds):
    for value, bits in bounds:
        assert 0 <= value < 2**bits


def build_condition(fact_registry_address: str, fact: bytes) -> int:
    """keccak(address, fact) & (2^250 - 1) (reference :14-21).  Needs web3, like the reference."""
    from web3 import Web3
    digest = Web3.solidityKeccak(["address", "bytes32"], [fact_registry_address, fact])
    return int.from_bytes(digest, "big") & (2**250 - 1)


# ------------------------------------------------------------------------------------ transfers
def _transfer_words(kind, sender_position_id, receiver_position_id, src_fee_position_id, nonce, amount, max_amount_fee,
                    expiration_timestamp):
    w0 = _pack([(sender_position_id, 64), (receiver_position_id, 64), (src_fee_position_id, 64), (nonce, 32)])
    w1 = _pack([(kind, 0), (amount, 64), (max_amount_fee, 64), (expiration_timestamp, 32)], pad_bits=81)
    return w0, w1


def get_conditional_transfer_msg_without_bounds(asset_id, asset_id_fee, receiver_public_key, condition, sender_position_id,
                                                receiver_position_id, src_fee_position_id, nonce, amount, max_amount_fee,
                                                expiration_timestamp, hash_function=pedersen_hash) -> int:
    w0, w1 = _transfer_words(CONDITIONAL_TRANSFER, sender_position_id, receiver_position_id, src_fee_position_id, nonce,
                             amount, max_amount_fee, expiration_timestamp)
    return _chain(hash_function, [asset_id, asset_id_fee, receiver_public_key, condition, w0, w1])


def get_conditional_transfer_msg(asset_id, asset_id_fee, receiver_public_key, condition, sender_position_id,
                                 receiver_position_id, src_fee_position_id, nonce, amount, max_amount_fee,
                                 expiration_timestamp, hash_function=pedersen_hash) -> int:
    _check([(amount, 64), (asset_id, 250), (asset_id_fee, 250), (condition, 251), (expiration_timestamp, 32),
            (src_fee_position_id, 64), (max_amount_fee, 64), (nonce, 32), (receiver_position_id, 64),
            (receiver_public_key, 251), (sender_position_id, 64)])
    return get_conditional_transfer_msg_without_bounds(
        asset_id, asset_id_fee, receiver_public_key, condition, sender_position_id, receiver_position_id,
        src_fee_position_id, nonce, amount, max_amount_fee, expiration_timestamp, hash_function=hash_function)


def get_transfer_msg_without_bounds(asset_id, asset_id_fee, receiver_public_key, sender_position_id, receiver_position_id,
                                    src_fee_position_id, nonce, amount, max_amount_fee, expiration_timestamp,
                                    hash_function=pedersen_hash) -> int:
    w0, w1 = _transfer_words(TRANSFER, sender_position_id, receiver_position_id, src_fee_position_id, nonce, amount,
                             max_amount_fee, expiration_timestamp)
    return _chain(hash_function, [asset_id, asset_id_fee, receiver_public_key, w0, w1])


def get_transfer_msg(asset_id, asset_id_fee, receiver_public_key, sender_position_id, receiver_position_id,
                     src_fee_position_id, nonce, amount, max_amount_fee, expiration_timestamp,
                     hash_function=pedersen_hash) -> int:
    _check([(amount, 64), (asset_id, 250), (asset_id_fee, 250), (expiration_timestamp, 32), (max_amount_fee, 64),
            (nonce, 32), (receiver_position_id, 64), (receiver_public_key, 251), (sender_position_id, 64),
            (src_fee_position_id, 64)])
    return get_transfer_msg_without_bounds(
        asset_id, asset_id_fee, receiver_public_key, sender_position_id, receiver_position_id, src_fee_position_id, nonce,
        amount, max_amount_fee, expiration_timestamp, hash_function=hash_function)


# ------------------------------------------------------------------------------------ withdrawal
def get_withdrawal_to_address_msg_without_bounds(asset_id_collateral, position_id, eth_address, nonce, expiration_timestamp,
                                                 amount, hash_function=pedersen_hash) -> int:
    word = _pack([(WITHDRAWAL_TO_ADDRESS, 0), (position_id, 64), (nonce, 32), (amount, 64), (expiration_timestamp, 32)],
                 pad_bits=49)
    return _chain(hash_function, [asset_id_collateral, int(eth_address, 16), word])


def get_withdrawal_to_address_msg(asset_id_collateral, position_id, eth_address, nonce, expiration_timestamp, amount,
                                  hash_function=pedersen_hash) -> int:
    _check([(asset_id_collateral, 250), (nonce, 32), (position_id, 64), (expiration_timestamp, 32), (amount, 64),
            (int(eth_address, 16), 160)])
    return get_withdrawal_to_address_msg_without_bounds(asset_id_collateral, position_id, eth_address, nonce,
                                                        expiration_timestamp, amount, hash_function=hash_function)


# ------------------------------------------------------------------------------------ limit orders
def limit_order_elements(asset_id_synthetic, asset_id_collateral, is_buying_synthetic, asset_id_fee, amount_synthetic,
                         amount_collateral, max_amount_fee, nonce, position_id, expiration_timestamp):
    """The five chain elements of a limit order (what k_pack_limit_orders builds on the device)."""
    sell, buy = (asset_id_synthetic, amount_synthetic), (asset_id_collateral, amount_collateral)
    if is_buying_synthetic:
        sell, buy = buy, sell
    w0 = _pack([(sell[1], 64), (buy[1], 64), (max_amount_fee, 64), (nonce, 32)])
    w1 = _pack([(LIMIT_ORDER_WITH_FEES, 0), (position_id, 64), (position_id, 64), (position_id, 64),
                (expiration_timestamp, 32)], pad_bits=17)
    return [sell[0], buy[0], asset_id_fee, w0, w1]


def get_limit_order_msg_without_bounds(asset_id_synthetic, asset_id_collateral, is_buying_synthetic, asset_id_fee,
                                       amount_synthetic, amount_collateral, max_amount_fee, nonce, position_id,
                                       expiration_timestamp, hash_function=pedersen_hash) -> int:
    return _chain(hash_function, limit_order_elements(asset_id_synthetic, asset_id_collateral, is_buying_synthetic,
                                                      asset_id_fee, amount_synthetic, amount_collateral, max_amount_fee,
                                                      nonce, position_id, expiration_timestamp))


def get_limit_order_msg(asset_id_synthetic, asset_id_collateral, is_buying_synthetic, asset_id_fee, amount_synthetic,
                        amount_collateral, max_amount_fee, nonce, position_id, expiration_timestamp,
                        hash_function=pedersen_hash) -> int:
    _check([(asset_id_synthetic, 128), (asset_id_collateral, 250), (asset_id_fee, 250), (amount_synthetic, 64),
            (amount_collateral, 64), (max_amount_fee, 64), (nonce, 32), (position_id, 64), (expiration_timestamp, 32)])
    return get_limit_order_msg_without_bounds(
        asset_id_synthetic, asset_id_collateral, is_buying_synthetic, asset_id_fee, amount_synthetic, amount_collateral,
        max_amount_fee, nonce, position_id, expiration_timestamp, hash_function=hash_function)


# ------------------------------------------------------------------------------------ oracle prices
def get_price_msg(oracle_name, asset_pair, timestamp, price, hash_function=pedersen_hash):
    _check([(oracle_name, 40), (asset_pair, 128), (timestamp, 32), (price, 120)])
    return hash_function(_pack([(asset_pair, 0), (oracle_name, 40)]), _pack([(price, 0), (timestamp, 32)]))


# ------------------------------------------------------------------------------------ device pipeline
ORDER_FIELDS = ("asset_id_synthetic", "asset_id_collateral", "is_buying_synthetic", "asset_id_fee", "amount_synthetic",
                "amount_collateral", "max_amount_fee", "nonce", "position_id", "expiration_timestamp")


def _order_arrays(orders):
    """orders: list of dicts (ORDER_FIELDS) or dict of equal-length sequences -> dict of numpy arrays in the
    layout of spg_limit_orders.  Bounds on the 64/32-bit fields are asserted here (the device ABI carries them
    at exactly that width); the three asset-id bounds are checked by the kernel."""
    from stark_perpetual_b200._lib import ints_to_limbs
    if not isinstance(orders, dict):
        orders = {f: [o[f] for o in orders] for f in ORDER_FIELDS}
    for f, bits in (("amount_synthetic", 64), ("amount_collateral", 64), ("max_amount_fee", 64), ("position_id", 64),
                    ("nonce", 32), ("expiration_timestamp", 32)):
        for v in orders[f]:
            assert 0 <= int(v) < 2**bits
    for f in ("asset_id_synthetic", "asset_id_collateral", "asset_id_fee"):
        for v in orders[f]:
            assert 0 <= int(v) < 2**256
    arr = {f: ints_to_limbs(orders[f]) for f in ("asset_id_synthetic", "asset_id_collateral", "asset_id_fee")}
    arr["is_buying_synthetic"] = np.array([1 if v else 0 for v in orders["is_buying_synthetic"]], dtype=np.uint8)
    for f in ("amount_synthetic", "amount_collateral", "max_amount_fee", "position_id"):
        arr[f] = np.array([int(v) for v in orders[f]], dtype=np.uint64)
    for f in ("nonce", "expiration_timestamp"):
        arr[f] = np.array([int(v) for v in orders[f]], dtype=np.uint32)
    return arr


def get_limit_order_msg_batch(orders):
    """[get_limit_order_msg(**o) for o in orders] in one device pipeline."""
    import stark_perpetual_b200 as spg
    from stark_perpetual_b200._lib import limbs_to_ints
    arr = _order_arrays(orders)
    out, st = spg.get_context(0).limit_order_msg(arr)
    assert not (st == 1).any()                                   # reference :226-230
    assert not (st == 2).any(), "Unhashable input."              # signature.py:313
    return limbs_to_ints(out)


def verify_limit_orders_batch(orders, rs, ss, public_keys):
    """[verify(get_limit_order_msg(**o), r, s, key)] with x-only keys; raises AssertionError where the reference
    would raise for any element."""
    import stark_perpetual_b200 as spg
    from stark_perpetual_b200._lib import ints_to_limbs
    arr = _order_arrays(orders)
    st = spg.get_context(0).limit_order_verify(arr, ints_to_limbs(rs), ints_to_limbs(ss), ints_to_limbs(public_keys))
    assert not (st == 2).any(), "precondition violated (order bound, unhashable message, or signature operand out of range)"
    return [bool(v) for v in st]


# the other messages through spg_message_hash_batch: (kind, full-width arguments, narrow arguments) in the order of
# include/spg.h; every bound the reference asserts is checked on the device
_MESSAGE_ARGS = {
    "transfer": (("asset_id", "asset_id_fee", "receiver_public_key"),
                 ("sender_position_id", "receiver_position_id", "src_fee_position_id", "nonce", "amount", "max_amount_fee",
                  "expiration_timestamp")),
    "conditional_transfer": (("asset_id", "asset_id_fee", "receiver_public_key", "condition"),
                             ("sender_position_id", "receiver_position_id", "src_fee_position_id", "nonce", "amount",
                              "max_amount_fee", "expiration_timestamp")),
    "withdrawal_to_address": (("asset_id_collateral", "eth_address"), ("position_id", "nonce", "amount", "expiration_timestamp")),
    "price": (("asset_pair", "price"), ("oracle_name", "timestamp")),
}


def get_msg_batch(kind, messages):
    """[get_<kind>_msg(**m) for m in messages] in one device pipeline (packing kernel + Pedersen chain); kind is one of
    'transfer', 'conditional_transfer', 'withdrawal_to_address', 'price'.  Raises AssertionError where the reference
    would for any element."""
    import stark_perpetual_b200 as spg
    from stark_perpetual_b200._lib import ints_to_limbs, limbs_to_ints
    fnames, inames = _MESSAGE_ARGS[kind]
    if not messages:
        return []

    def val(m, name):
        v = m[name]
        return int(v, 16) if isinstance(v, str) else int(v)
    felts, ints = [], []
    for name in fnames:
        col = [val(m, name) for m in messages]
        for v in col:
            assert 0 <= v < 2**256
        felts.append(ints_to_limbs(col))
    for name in inames:
        col = [val(m, name) for m in messages]
        for v in col:
            assert 0 <= v < 2**64
        ints.append(np.array(col, dtype=np.uint64))
    out, st = spg.get_context(0).message_hash(kind, felts, ints)
    assert not (st == 1).any()                                   # the bounds of reference :38-48, :110-119, :174-179, :314-317
    assert not (st == 2).any(), "Unhashable input."              # signature.py:313
    return limbs_to_ints(out)


def get_transfer_msg_batch(messages):
    return get_msg_batch("transfer", messages)


def get_conditional_transfer_msg_batch(messages):
    return get_msg_batch("conditional_transfer", messages)


def get_withdrawal_to_address_msg_batch(messages):
    return get_msg_batch("withdrawal_to_address", messages)


def get_price_msg_batch(messages):
    return get_msg_batch("price", messages)
