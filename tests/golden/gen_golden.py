#!/usr/bin/env python3
"""Generate tests/golden/crypto_golden.json by RUNNING THE REFERENCE's own Python
(/root/reference/src, unmodified) in this container.  Two harness-side import shims are
needed because the image lacks `ecdsa`/`web3`/`mypy_extensions` and ships a newer sympy
(SURVEY.md section 0.5): `sympy.core.numbers.igcdex` is aliased from sympy.core.intfunc, and
`ecdsa.rfc6979.generate_k` is provided by oracle.ecdsa.rfc6979_generate_k (which the JS
KATs at signature.spec.js:96-137 pin).  The GPU box has no /root/reference: tests read only
the JSON this script wrote.

Run:  python tests/golden/gen_golden.py        (about a minute)
"""
import json
import os
import random
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "src"))

import sympy.core.numbers  # noqa: E402
import sympy.core.intfunc  # noqa: E402
sympy.core.numbers.igcdex = sympy.core.intfunc.igcdex

from oracle import ecdsa as o_ecdsa  # noqa: E402

m = types.ModuleType("ecdsa"); m.rfc6979 = types.ModuleType("ecdsa.rfc6979")
m.rfc6979.generate_k = o_ecdsa.rfc6979_generate_k
sys.modules["ecdsa"] = m; sys.modules["ecdsa.rfc6979"] = m.rfc6979
me = types.ModuleType("mypy_extensions"); me.VarArg = lambda t: t
sys.modules["mypy_extensions"] = me
w3 = types.ModuleType("web3"); w3.Web3 = object
sys.modules["web3"] = w3

from starkware.crypto.signature import signature as ref  # noqa: E402
from services.perpetual.public import perpetual_messages as ref_msg  # noqa: E402

P = ref.FIELD_PRIME
rng = random.Random(20261017)
H = hex


def rand_felt():
    return rng.randrange(P)


def main():
    out = {"_generated_by": "tests/golden/gen_golden.py from /root/reference (commit 40f02826)"}

    # ---- Pedersen: reference KATs + edge set + random (signature.py:296-318) -----------------
    td = json.load(open(os.path.join(
        REF, "src/starkware/crypto/signature/test/config/signature_test_data.json")))
    ped = []
    for k in ("pedersen_hash_data_1", "pedersen_hash_data_2"):
        d = td["hash_test"][k]
        a, b, o = int(d["input_1"], 16), int(d["input_2"], 16), int(d["output"], 16)
        assert ref.pedersen_hash(a, b) == o
        ped.append([H(a), H(b), H(o), "kat"])
    edge = [0, 1, 2, P - 1, P - 2, 2**248 - 1, 2**248, 2**251, 2**251 - 1, 2**250 + 12345, (1 << 192) - 1]
    for a in edge:
        for b in (0, 1, P - 1, rand_felt()):
            ped.append([H(a), H(b), H(ref.pedersen_hash(a, b)), "edge"])
    for _ in range(96):
        a, b = rand_felt(), rand_felt()
        ped.append([H(a), H(b), H(ref.pedersen_hash(a, b)), "random"])
    out["pedersen"] = ped
    # single-element and out-of-range behaviour
    out["pedersen_single"] = [[H(a), H(ref.pedersen_hash(a))] for a in (0, 1, P - 1, rand_felt())]

    # ---- priv -> pub (keys_precomputed.json) -------------------------------------------------
    keys = json.load(open(os.path.join(
        REF, "src/starkware/crypto/signature/src/config/keys_precomputed.json")))
    kp = []
    for priv, pub in list(keys.items()):
        if priv.startswith("_"):
            continue
        assert ref.private_to_stark_key(int(priv, 16)) == int(pub, 16)
        kp.append([priv, pub])
    out["keys"] = kp

    # ---- verify: positives / negatives from signature_test_data.json -------------------------
    ver = []

    def add_verify(msg, r, s, pub, tag):
        try:
            res = ref.verify(msg, r, s, pub)
            res = 1 if res else 0
        except AssertionError:
            res = 2
        if isinstance(pub, tuple):
            ver.append([H(msg), H(r), H(s), [H(pub[0]), H(pub[1])], res, tag])
        else:
            ver.append([H(msg), H(r), H(s), H(pub), res, tag])

    md = td["meta_data"]
    for name, node in (("party_a_order", td["settlement"]["party_a_order"]),
                       ("party_b_order", td["settlement"]["party_b_order"]),
                       ("transfer_order", td["transfer_order"]),
                       ("conditional_transfer_order", td["conditional_transfer_order"]),
                       ("transfer_order_2nd_valid_range", td["transfer_order_2nd_valid_range"]),
                       ("order_with_vault_id_in_2nd_range", td["order_with_vault_id_in_2nd_range"])):
        msg = int(md[name]["message_hash"], 16)
        r, s = int(node["signature"]["r"], 16), int(node["signature"]["s"], 16)
        add_verify(msg, r, s, int(node["public_key"], 16), "testdata:" + name)
    mo = td["multi_asset_order"]
    pub = ref.private_to_stark_key(int(md["multi_asset_order"]["private_key"], 16))
    add_verify(int(md["multi_asset_order"]["message_hash"], 16), int(mo["signature"]["r"], 16),
               int(mo["signature"]["s"], 16), pub, "testdata:multi_asset_order")

    # random signatures made by the REFERENCE sign (with the RFC 6979 shim), plus corruptions
    sigs = []
    for i in range(14):
        priv = rng.randrange(1, ref.EC_ORDER)
        msg = rng.randrange(2**251)
        r, s = ref.sign(msg, priv)
        pubpt = ref.private_key_to_ec_point_on_stark_curve(priv)
        sigs.append([H(msg), H(priv), H(r), H(s)])
        add_verify(msg, r, s, pubpt[0], "valid:x-only")
        if i % 2 == 0:
            add_verify(msg, r, s, pubpt, "valid:point")
            add_verify(msg, r, s, (pubpt[0], P - pubpt[1]), "wrong-y:point")
        kind = i % 7
        if kind == 0:
            add_verify(msg ^ 1, r, s, pubpt[0], "bad:msg")
        elif kind == 1:
            add_verify(msg, r ^ 2, s, pubpt[0], "bad:r")
        elif kind == 2:
            add_verify(msg, r, s ^ 4, pubpt[0], "bad:s")
        elif kind == 3:
            add_verify(msg, r, s, pubpt[0] ^ 1, "bad:key")
        elif kind == 4:
            add_verify(0, r, s, pubpt[0], "msg=0")
        elif kind == 5:
            add_verify(msg, 0, s, pubpt[0], "raise:r=0")
            add_verify(msg, r, 0, pubpt[0], "raise:s=0")
            add_verify(msg, 2**251, s, pubpt[0], "raise:r=2^251")
            add_verify(2**251, r, s, pubpt[0], "raise:msg=2^251")
            add_verify(msg, r, ref.EC_ORDER, pubpt[0], "raise:s=n")
        else:
            add_verify(msg, r, s, (pubpt[0], (pubpt[1] + 1) % P), "raise:off-curve")
    # the reachable structural rejections: zG == +-rQ and Q == shift point
    z = rng.randrange(1, 2**251)
    k = rng.randrange(1, ref.EC_ORDER)
    r = ref.ec_mult(k, ref.EC_GEN, ref.ALPHA, P)[0]
    d = z * pow(r, -1, ref.EC_ORDER) % ref.EC_ORDER
    s = 2 * z * pow(k, -1, ref.EC_ORDER) % ref.EC_ORDER
    if r < 2**251 and 1 <= pow(s, -1, ref.EC_ORDER) < 2**251:
        add_verify(z, r, s, ref.private_key_to_ec_point_on_stark_curve(d), "structural:zG==rQ")
    add_verify(z, r, s, tuple(ref.SHIFT_POINT), "structural:Q==shift")
    add_verify(z, r, s, tuple(ref.EC_GEN), "structural:Q==G")
    # crafted public keys Q = (+-2^k - (r mod 2^k))^-1 * shift: the step-k partial sum of r*Q + shift
    # collides in x with 2^k * Q (signature.py:183), so the reference answers False at step k
    for kk, sign_ in ((0, 1), (1, 1), (7, -1), (100, 1), (250, -1)):
        rr = rng.randrange(2**250, 2**251)
        c = pow((sign_ * 2**kk - (rr % 2**kk)) % ref.EC_ORDER, -1, ref.EC_ORDER)
        Q = ref.ec_mult(c, tuple(ref.SHIFT_POINT), ref.ALPHA, P)
        ss = rng.randrange(1, ref.EC_ORDER)
        while not (1 <= pow(ss, -1, ref.EC_ORDER) < 2**251):
            ss = rng.randrange(1, ref.EC_ORDER)
        add_verify(z, rr, ss, Q, "structural:collision@%d" % kk)
    out["verify"] = ver
    out["sign"] = sigs

    # JS deterministic-sign KATs (signature.spec.js:96-137) replayed through the reference sign
    priv = 0x2dccce1da22003777062ee0870e9881b460a8b7eca276870f57c601f182136c
    js = []
    for mh, er, es in (
        ("c465dd6b1bbffdb05442eb17f5ca38ad1aa78a6f56bf4415bdee219114a47",
         "5f496f6f210b5810b2711c74c15c05244dad43d18ecbbdbe6ed55584bc3b0a2",
         "4e8657b153787f741a67c0666bad6426c3741b478c8eaa3155196fc571416f3"),
        ("00c465dd6b1bbffdb05442eb17f5ca38ad1aa78a6f56bf4415bdee219114a47",
         "5f496f6f210b5810b2711c74c15c05244dad43d18ecbbdbe6ed55584bc3b0a2",
         "4e8657b153787f741a67c0666bad6426c3741b478c8eaa3155196fc571416f3"),
        ("c465dd6b1bbffdb05442eb17f5ca38ad1aa78a6f56bf4415bdee219114a47a",
         "233b88c4578f0807b4a7480c8076eca5cfefa29980dd8e2af3c46a253490e9c",
         "28b055e825bc507349edfb944740a35c6f22d377443c34742c04e0d82278cf1"),
        ("7465dd6b1bbffdb05442eb17f5ca38ad1aa78a6f56bf4415bdee219114a47a1",
         "b6bee8010f96a723f6de06b5fa06e820418712439c93850dd4e9bde43ddf",
         "1a3d2bc954ed77e22986f507d68d18115fa543d1901f5b4620db98e2f6efd80")):
        r, s = ref.sign(int(mh, 16), priv)
        assert (r, s) == (int(er, 16), int(es, 16)), "RFC 6979 shim does not hit the JS KAT"
        js.append([mh, H(priv), er, es])
    out["sign_js_kat"] = js

    # get_y_coordinate / is_valid_stark_key
    ys = []
    for _ in range(12):
        x = rand_felt()
        try:
            ys.append([H(x), H(ref.get_y_coordinate(x))])
        except ref.InvalidPublicKeyError:
            ys.append([H(x), None])
    out["get_y"] = ys
    out["grind_key"] = [[H(a), H(b), H(ref.grind_key(a, b))] for a, b in
                        ((0x86F3E7293141F20A8BAFF320E8EE4ACCB9D4A4BF2B4D295E8CEE784DB46E0519, ref.EC_ORDER),
                         (rand_felt(), ref.EC_ORDER), (rand_felt(), 2**200 + 7))]

    # ---- perpetual message hashes (perpetual_messages_test.py + random) ----------------------
    pre = json.load(open(os.path.join(
        REF, "src/services/perpetual/public/perpetual_messages_precomputed.json")))
    out["messages_precomputed"] = {k: v for k, v in pre.items() if not k.startswith("_")}
    msgs = []
    for _ in range(8):
        a = dict(asset_id_synthetic=rng.randrange(2**128), asset_id_collateral=rng.randrange(2**250),
                 is_buying_synthetic=rng.randrange(2), asset_id_fee=rng.randrange(2**250),
                 amount_synthetic=rng.randrange(2**64), amount_collateral=rng.randrange(2**64),
                 max_amount_fee=rng.randrange(2**64), nonce=rng.randrange(2**32),
                 position_id=rng.randrange(2**64), expiration_timestamp=rng.randrange(2**32))
        msgs.append(["limit_order", a, H(ref_msg.get_limit_order_msg(**a))])
        t = dict(asset_id=rng.randrange(2**250), asset_id_fee=rng.randrange(2**250),
                 receiver_public_key=rng.randrange(2**251), sender_position_id=rng.randrange(2**64),
                 receiver_position_id=rng.randrange(2**64), src_fee_position_id=rng.randrange(2**64),
                 nonce=rng.randrange(2**32), amount=rng.randrange(2**64),
                 max_amount_fee=rng.randrange(2**64), expiration_timestamp=rng.randrange(2**32))
        msgs.append(["transfer", t, H(ref_msg.get_transfer_msg(**t))])
        c = dict(t, condition=rng.randrange(2**251))
        msgs.append(["conditional_transfer", c, H(ref_msg.get_conditional_transfer_msg(**c))])
        w = dict(asset_id_collateral=rng.randrange(2**250), position_id=rng.randrange(2**64),
                 eth_address=H(rng.randrange(2**160)), nonce=rng.randrange(2**32),
                 expiration_timestamp=rng.randrange(2**32), amount=rng.randrange(2**64))
        msgs.append(["withdrawal_to_address", w, H(ref_msg.get_withdrawal_to_address_msg(**w))])
        p = dict(oracle_name=rng.randrange(2**40), asset_pair=rng.randrange(2**128),
                 timestamp=rng.randrange(2**32), price=rng.randrange(2**120))
        msgs.append(["price", p, H(ref_msg.get_price_msg(**p))])
    out["messages"] = msgs

    with open(os.path.join(HERE, "crypto_golden.json"), "w") as f:
        json.dump(out, f, indent=0, default=str)
    print("wrote crypto_golden.json:", {k: len(v) for k, v in out.items() if not k.startswith("_")})


if __name__ == "__main__":
    main()
