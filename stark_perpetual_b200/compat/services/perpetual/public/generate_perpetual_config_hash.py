#!/usr/bin/env python3
"""GPU-backed mirror of the reference's `services.perpetual.public.generate_perpetual_config_hash`
(src/services/perpetual/public/generate_perpetual_config_hash.py): the general-config hash and one hash per synthetic
asset, each a left-folded Pedersen chain  h = 0; h = H(h, v) for v in fields + [len(fields)]  (:127-130, :169-172) -- what
`hash_init / hash_update_single / hash_finalize` compute on the Cairo side
(src/services/perpetual/cairo/definitions/general_config_hash.cairo:97-140).

Same names, argument meaning, assertions and output text as the reference.  What differs is the schedule: the field lists
are built first and all chains of one length are hashed by ONE call of spg_pedersen_chain_batch, so a config with hundreds
of synthetic assets costs a handful of launches instead of thousands of single hashes.
"""
import argparse
import sys

import numpy as np

import stark_perpetual_b200 as _spg
from services.perpetual.definitions.general_config import GENERAL_CONFIG_HASH_VERSION
from services.perpetual.public.definitions.constants import ASSET_ID_UPPER_BOUND, RISK_UPPER_BOUND
from stark_perpetual_b200._lib import FIELD_PRIME, ints_to_limbs, limbs_to_ints

CONFIG_FILE_NAME = "production_general_config.yml"
HASH_BYTES = 32
ASSET_ID_BYTES = 15
assert 2 ** (ASSET_ID_BYTES * 8) == ASSET_ID_UPPER_BOUND


def convert2int(val) -> int:
    """decimal string, hex string, bool or int -> int   (:42-53)"""
    if type(val) in (int, bool):
        return int(val)
    assert type(val) is str, "Unsupported type."
    return int(val, 16) if len(val) > 2 and val[:2] == "0x" else int(val, 10)


def bytes2str(val: bytes) -> str:
    return "0x" + val.hex()


def pad_hex_string(val: str, bytes_len: int) -> str:
    assert val[:2] == "0x"
    digits = val[2:]
    assert len(digits) <= 2 * bytes_len
    return "0x" + digits.rjust(2 * bytes_len, "0")


def _get(d: dict, *path):
    for key in path:
        assert key in d
        d = d[key]
    return d


# the general config's hashed fields, in order (:111-124)
_GENERAL_FIELDS = (("max_funding_rate",), ("collateral_asset_info", "asset_id"), ("collateral_asset_info", "resolution"),
                   ("fee_position_info", "position_id"), ("fee_position_info", "public_key"), ("positions_tree_height",),
                   ("orders_tree_height",), ("timestamp_validation_config", "price_validity_period"),
                   ("timestamp_validation_config", "funding_validity_period"), ("data_availability_mode",),
                   ("is_risk_by_balance_only",))


def general_config_fields(config: dict) -> list:
    vals = [GENERAL_CONFIG_HASH_VERSION] + [_get(config, *path) for path in _GENERAL_FIELDS]
    return [convert2int(v) for v in vals] + [len(vals)]


def asset_fields(config: dict, asset_id: str) -> list:
    info = _get(config, "synthetic_assets_info", asset_id)
    segments = _get(info, "risk_factor", "segments")
    signed_ids, signers = _get(info, "oracle_price_signed_asset_ids"), _get(info, "oracle_price_signers")
    vals = [asset_id, _get(info, "resolution"), len(segments)]
    vals += [seg["upper_bound"] * RISK_UPPER_BOUND + int(seg["risk"]) for seg in segments]          # (:161-162)
    vals += [len(signed_ids)] + list(signed_ids) + [_get(info, "oracle_price_quorum"), len(signers)] + list(signers)
    return [convert2int(v) for v in vals] + [len(vals)]


def hash_chains(field_lists) -> list:
    """left-folded chains from 0, batched per length on the device -> 32-byte big-endian digests"""
    ctx = _spg.get_context(0)
    out = [None] * len(field_lists)
    by_len = {}
    for k, f in enumerate(field_lists):
        assert all(0 <= v < 2 ** 256 for v in f)                      # to_bytes' range (utils.py:414-451)
        assert all(v < FIELD_PRIME for v in f)                        # the hash's own assertion (signature.py:307)
        by_len.setdefault(len(f), []).append(k)
    for ln, idx in by_len.items():
        elems = ints_to_limbs([v for k in idx for v in [0] + field_lists[k]])
        res, st = ctx.pedersen_chain(elems, ln + 1)
        assert not (st == 2).any(), "Unhashable input."
        assert not st.any()
        for k, h in zip(idx, limbs_to_ints(res)):
            out[k] = h.to_bytes(HASH_BYTES, "big")
    return out


def calculate_general_config_hash(config: dict) -> bytes:
    return hash_chains([general_config_fields(config)])[0]


def calculate_asset_hash(config: dict, asset_id: str) -> bytes:
    return hash_chains([asset_fields(config, asset_id)])[0]


def generate_config_hashes(config: dict) -> str:
    assets = list(config["synthetic_assets_info"].keys())
    digests = hash_chains([general_config_fields(config)] + [asset_fields(config, a) for a in assets])
    lines = ["Global config hash: %s\n" % bytes2str(digests[0])]
    lines += ["asset_id: %s, config_hash: %s\n" % (pad_hex_string(a, ASSET_ID_BYTES), bytes2str(d)) for a, d in zip(assets, digests[1:])]
    return "".join(lines) + "\n"


def parse_cmdline():
    parser = argparse.ArgumentParser(description="Calculates dYdX general config and synthetic asset hash values.")
    parser.add_argument("--general_config_file_name", type=str, default=CONFIG_FILE_NAME,
                        help="Input YAML file containing the general configuration.")
    return parser.parse_args()


def main():
    import yaml
    args = parse_cmdline()
    with open(args.general_config_file_name, "r") as f:
        config = yaml.load(f, Loader=yaml.FullLoader)
    print(generate_config_hashes(config))


assert np is not None
if __name__ == "__main__":
    sys.exit(main())
