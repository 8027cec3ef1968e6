#!/usr/bin/env python3
"""Golden vectors for the config-hash chains (missing item 3): runs the REFERENCE's own
generate_perpetual_config_hash.py (imported from /root/reference/src) on a synthetic general config.

Two of its imports are not part of the published tree and are stubbed with the values the Cairo program itself uses
(general_config_hash.cairo:101-102, constants.cairo:11, :26, :42); its hash, fast_pedersen_hash.pedersen_hash_func, needs
fastecdsa (absent here) and is stubbed with the reference's pure-Python signature.pedersen_hash behind the same byte ABI
(fast_pedersen_hash.py:47-52).  Everything else -- field order, conversions, the length suffix, the output text -- is the
reference's code.  Writes tests/golden/config_hash_golden.json."""
import json
import os
import random
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
from oracle import refenv  # noqa: E402

_sig, _mu, _pm = refenv.import_reference()     # sympy / ecdsa / web3 shims + /root/reference/src on sys.path
pedersen_hash = _sig.pedersen_hash             # the reference's

gc = types.ModuleType("services.perpetual.definitions.general_config")
gc.GENERAL_CONFIG_HASH_VERSION = int.from_bytes(b"PerpetualConfig1", "big")
sys.modules["services.perpetual.definitions"] = types.ModuleType("services.perpetual.definitions")
sys.modules["services.perpetual.definitions.general_config"] = gc
cs = types.ModuleType("services.perpetual.public.definitions.constants")
cs.ASSET_ID_UPPER_BOUND, cs.RISK_UPPER_BOUND = 2 ** 120, 2 ** 32
sys.modules["services.perpetual.public.definitions"] = types.ModuleType("services.perpetual.public.definitions")
sys.modules["services.perpetual.public.definitions.constants"] = cs
fp = types.ModuleType("starkware.crypto.signature.fast_pedersen_hash")
fp.pedersen_hash_func = lambda x, y: pedersen_hash(int.from_bytes(x, "big"), int.from_bytes(y, "big")).to_bytes(32, "big")
sys.modules["starkware.crypto.signature.fast_pedersen_hash"] = fp

from services.perpetual.public import generate_perpetual_config_hash as ref  # noqa: E402


def make_config(seed, n_assets):
    rng = random.Random(seed)

    def asset(k):
        name = ("SYN%d-%d" % (k, rng.randrange(3, 12))).encode().ljust(15, b"\0")
        n_seg, n_ids, n_signers = rng.randrange(1, 4), rng.randrange(1, 4), rng.randrange(1, 6)
        bound = 0
        segs = []
        for _ in range(n_seg):
            bound += rng.randrange(1, 2 ** 40)
            segs.append({"upper_bound": bound, "risk": str(rng.randrange(1, 2 ** 32)) if rng.random() < 0.5 else rng.randrange(1, 2 ** 32)})
        return "0x" + name.hex(), {
            "resolution": rng.choice([10 ** rng.randrange(3, 10), hex(10 ** rng.randrange(3, 10)), str(10 ** rng.randrange(3, 10))]),
            "risk_factor": {"segments": segs},
            "oracle_price_signed_asset_ids": [hex(rng.randrange(2 ** 128)) for _ in range(n_ids)],
            "oracle_price_quorum": rng.randrange(1, n_signers + 1),
            "oracle_price_signers": [hex(rng.randrange(2 ** 250)) for _ in range(n_signers)],
        }
    return {
        "max_funding_rate": rng.randrange(1, 2 ** 32),
        "collateral_asset_info": {"asset_id": hex(rng.randrange(2 ** 250)), "resolution": 10 ** 6},
        "fee_position_info": {"position_id": str(rng.randrange(2 ** 64)), "public_key": hex(rng.randrange(2 ** 250))},
        "positions_tree_height": 64, "orders_tree_height": 64,
        "timestamp_validation_config": {"price_validity_period": 31 * 24 * 3600, "funding_validity_period": "604800"},
        "data_availability_mode": rng.randrange(2), "is_risk_by_balance_only": bool(rng.randrange(2)),
        "synthetic_assets_info": dict(asset(k) for k in range(n_assets)),
    }


def main():
    cases = []
    for seed, n_assets in ((1, 3), (2, 6)):
        cfg = make_config(seed, n_assets)
        cases.append({"config": cfg, "output": ref.generate_config_hashes(cfg),
                      "general": "0x" + ref.calculate_general_config_hash(cfg).hex(),
                      "assets": {a: "0x" + ref.calculate_asset_hash(cfg, a).hex() for a in cfg["synthetic_assets_info"]}})
    with open(os.path.join(HERE, "config_hash_golden.json"), "w") as f:
        json.dump({"source": "src/services/perpetual/public/generate_perpetual_config_hash.py run by tests/golden/gen_golden3.py",
                   "cases": cases}, f, indent=1)
    print("wrote %d cases" % len(cases))


if __name__ == "__main__":
    main()
