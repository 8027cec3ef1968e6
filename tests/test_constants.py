"""CPU: row a1 -- the constant tables this repo derives (tools/gen_curve_params.py -> data/curve_params.json ->
csrc/curve_params.inc, expanded by doubling chains) equal the reference's pedersen_params.json
(signature.py:38-68, nothing_up_my_sleeve_gen.py:35-104).  Always checked against a digest of the reference's table taken
in the build container; checked element by element whenever the reference file itself is reachable (/root/reference or
the staged oracle/_ref)."""
import hashlib
import json
import os
import re

from oracle import params, refenv

# sha256 over the 506 points of /root/reference/src/starkware/crypto/signature/pedersen_params.json (commit 40f02826),
# each coordinate as 32 big-endian bytes, x before y
REFERENCE_TABLE_SHA256 = "58d83b9f87ef1e7e49924404f21a2e78dad61aa194eb9fcf76a75ff7f6cc6795"
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _digest(points):
    return hashlib.sha256(b"".join(int(c).to_bytes(32, "big") for pt in points for c in pt)).hexdigest()


def test_constant_points_equal_reference_table():
    assert len(params.CONSTANT_POINTS) == 506
    assert _digest(params.CONSTANT_POINTS) == REFERENCE_TABLE_SHA256
    assert params.FIELD_PRIME == 3618502788666131213697322783095070105623107215331596699973092056135872020481
    assert params.EC_ORDER == 3618502788666131213697322783095070105526743751716087489154079457884512865583
    assert (params.ALPHA, params.FIELD_GEN) == (1, 3)
    assert params.BETA == 3141592653589793238462643383279502884197169399375105820974944592307816406665
    src = refenv.ref_src()
    if src is not None:
        ref = json.load(open(os.path.join(src, "starkware", "crypto", "signature", "pedersen_params.json")))
        assert [list(p) for p in params.CONSTANT_POINTS] == ref["CONSTANT_POINTS"]
        for k in ("FIELD_PRIME", "FIELD_GEN", "EC_ORDER", "ALPHA", "BETA"):
            assert getattr(params, k) == ref[k], k


def test_device_table_source_matches_json():
    """csrc/curve_params.inc (what the kernels' tables are expanded from) holds the same base points as the JSON."""
    inc = open(os.path.join(ROOT, "stark_perpetual_b200", "csrc", "curve_params.inc")).read()
    words = [int(w, 16) for w in re.findall(r"0x([0-9a-fA-F]{16})ULL", inc)]
    prm = json.load(open(os.path.join(ROOT, "stark_perpetual_b200", "data", "curve_params.json")))
    felts = {sum(words[i + k] << (64 * k) for k in range(4)) for i in range(0, len(words) - 3)}
    for name, (x, y) in prm["BASE_POINTS"].items():
        assert int(x, 16) in felts and int(y, 16) in felts, name
    assert int(prm["BETA"], 16) in felts
