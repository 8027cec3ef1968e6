"""`services.perpetual.definitions.general_config` is imported by the reference's config-hash script
(src/services/perpetual/public/generate_perpetual_config_hash.py:31) but is not part of the published tree.  The one name
the script needs is restated from the Cairo source that consumes the hash
(src/services/perpetual/cairo/definitions/general_config_hash.cairo:101-102: the felt of the string "PerpetualConfig1")."""
GENERAL_CONFIG_HASH_VERSION = int.from_bytes(b"PerpetualConfig1", "big")
assert GENERAL_CONFIG_HASH_VERSION == 106864982745153081011865306738524251953
