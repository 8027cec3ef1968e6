#!/usr/bin/env python3
"""Stage the reference files of THIS path into oracle/_ref/src (git-ignored; not gpurun-ignored, so the copy travels to
the GPU box where /root/reference does not exist).  Run by __graft_entry__.build() when /root/reference is present.

Staged, unmodified, in the reference's own directory layout -- plus the three JSON fixtures placed next to
stark_cli_test.py, where that test opens them (its Bazel rule copies them there; `os.path.dirname(__file__)`,
stark_cli_test.py:28-36):
  * the CPU baseline of the crypto rows (bench.py cpu_baseline.kind = "reference"): signature.py, math_utils.py,
    fast_pedersen_hash.py, pedersen_params.json, perpetual_messages.py;
  * the reference's own tests for this path, run UNCHANGED against the GPU-backed compat tree by
    tests/test_gpu_reference_tests.py: perpetual_messages_test.py, stark_cli_test.py, stark_cli.py and their fixtures;
  * starkware/python/merkle_tree.py, the hint helper that pins the update-tree shape of the state-tree row (f-4).
Nothing is copied into the repository's history.
"""
import os
import shutil
import sys

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "src")

FILES = [
    "starkware/__init__.py", "starkware/crypto/__init__.py", "starkware/crypto/signature/__init__.py",
    "starkware/crypto/signature/signature.py", "starkware/crypto/signature/math_utils.py",
    "starkware/crypto/signature/fast_pedersen_hash.py", "starkware/crypto/signature/pedersen_params.json",
    "starkware/python/__init__.py", "starkware/python/merkle_tree.py", "starkware/python/math_utils.py",
    "starkware/python/utils.py", "starkware/python/utils_stub_module.py",
    "services/__init__.py", "services/perpetual/__init__.py", "services/perpetual/public/__init__.py",
    "services/perpetual/public/perpetual_messages.py", "services/perpetual/public/perpetual_messages_test.py",
    "services/perpetual/public/perpetual_messages_precomputed.json",
    "services/perpetual/public/stark_cli.py", "services/perpetual/public/stark_cli_test.py",
    # the config-hash script: run unchanged (by path) against compat's hash and the two definitions modules it imports
    "services/perpetual/public/generate_perpetual_config_hash.py",
    # the program-hash test (row f-2): run unchanged against compat's hash_program on a synthetic compiled program
    "starkware/cairo/__init__.py", "starkware/cairo/bootloaders/__init__.py",
    "starkware/cairo/bootloaders/program_hash_test_utils.py",
    "services/perpetual/cairo/__init__.py", "services/perpetual/cairo/program_hash_test.py",
    "services/perpetual/cairo/program_hash.json",
]
# fixtures the Bazel test rule places beside stark_cli_test.py
BESIDE_CLI_TEST = [
    "starkware/crypto/signature/test/config/signature_test_data.json",
    "starkware/crypto/signature/src/config/keys_precomputed.json",
]


def stage():
    if not os.path.isdir(REF):
        print("stage_ref: %s not present, nothing staged" % REF)
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(src):
            shutil.copyfile(src, dst)
        elif rel.endswith("__init__.py"):
            continue                          # a namespace directory in the reference: stays one in the copy
        else:
            raise FileNotFoundError(src)
    for rel in BESIDE_CLI_TEST:
        shutil.copyfile(os.path.join(REF, rel), os.path.join(DST, "services/perpetual/public", os.path.basename(rel)))
    print("stage_ref: staged %d files into %s" % (len(FILES) + len(BESIDE_CLI_TEST), DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
