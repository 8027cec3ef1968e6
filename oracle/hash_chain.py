"""compute_hash_chain / compute_program_hash_chain, plain Python.  TEST INFRASTRUCTURE.

The algorithm lives in cairo-lang, an UN-VENDORED dependency of the reference (`cairo-lang==0.0.0+local`,
scripts/requirements-gen.txt:2; call site src/starkware/cairo/bootloaders/program_hash_test_utils.py:3-9); restated here
from its published definition.  PARITY UNPINNED beyond the hash function: the reference's golden program hashes
(src/services/perpetual/cairo/program_hash.json:2) need the Cairo compiler's output, which cannot be produced here.
"""
from .pedersen import pedersen_hash


def compute_hash_chain(data, hash_func=pedersen_hash):
    assert len(data) >= 1
    h = data[-1]
    for x in reversed(data[:-1]):
        h = hash_func(x, h)
    return h


def compute_program_hash_chain(builtins, main, data, bootloader_version=0, hash_func=pedersen_hash):
    header = [bootloader_version, main, len(builtins)] + [int.from_bytes(b.encode("ascii"), "big") for b in builtins]
    chain = header + list(data)
    return compute_hash_chain([len(chain)] + chain, hash_func)
