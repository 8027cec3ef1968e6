"""General-config and synthetic-asset hashes, restated on the CPU.  TEST INFRASTRUCTURE.

Follows src/services/perpetual/public/generate_perpetual_config_hash.py:73-131 (general config) and :134-173 (one synthetic
asset): the listed fields, then their count, folded from 0 with the Pedersen hash.  Pinned by
tests/golden/config_hash_golden.json, which tests/golden/gen_golden3.py produced by running the reference's own functions."""
from .pedersen import pedersen_hash

GENERAL_CONFIG_HASH_VERSION = int.from_bytes(b"PerpetualConfig1", "big")      # general_config_hash.cairo:101-102
RISK_UPPER_BOUND = 2 ** 32                                                     # constants.cairo:26, :42


def _int(v):
    if isinstance(v, (int, bool)):
        return int(v)
    return int(v, 16) if v[:2] == "0x" and len(v) > 2 else int(v)


def _fold(values):
    h = 0
    for v in values + [len(values)]:
        h = pedersen_hash(h, _int(v))
    return h


def general_config_hash(cfg):
    return _fold([GENERAL_CONFIG_HASH_VERSION, cfg["max_funding_rate"], cfg["collateral_asset_info"]["asset_id"],
                  cfg["collateral_asset_info"]["resolution"], cfg["fee_position_info"]["position_id"],
                  cfg["fee_position_info"]["public_key"], cfg["positions_tree_height"], cfg["orders_tree_height"],
                  cfg["timestamp_validation_config"]["price_validity_period"],
                  cfg["timestamp_validation_config"]["funding_validity_period"], cfg["data_availability_mode"],
                  cfg["is_risk_by_balance_only"]])


def asset_hash(cfg, asset_id):
    a = cfg["synthetic_assets_info"][asset_id]
    seg = a["risk_factor"]["segments"]
    vals = [asset_id, a["resolution"], len(seg)] + [s["upper_bound"] * RISK_UPPER_BOUND + int(s["risk"]) for s in seg]
    vals += [len(a["oracle_price_signed_asset_ids"])] + list(a["oracle_price_signed_asset_ids"])
    vals += [a["oracle_price_quorum"], len(a["oracle_price_signers"])] + list(a["oracle_price_signers"])
    return _fold(vals)
