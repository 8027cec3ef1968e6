"""`services.perpetual.public.definitions.constants` is imported by the reference's config-hash script
(src/services/perpetual/public/generate_perpetual_config_hash.py:32) but is not part of the published tree.  The two bounds
it needs are restated from the Cairo constants the program itself uses
(src/services/perpetual/cairo/definitions/constants.cairo:11 ASSET_ID_UPPER_BOUND, :42 RISK_UPPER_BOUND = FXP_32_ONE)."""
ASSET_ID_UPPER_BOUND = 2 ** 120
FXP_32_ONE = 2 ** 32
RISK_UPPER_BOUND = FXP_32_ONE
