"""compute_hash_chain of cairo-lang (starkware/cairo/common/hash_chain.py -- an UN-VENDORED dependency of the reference,
`cairo-lang==0.0.0+local`, scripts/requirements-gen.txt:2), restated from its published definition and backed by libspg:

    h(data[0], h(data[1], h(..., h(data[n-2], data[n-1]))))

The reference reaches it through compute_program_hash_chain (src/starkware/cairo/bootloaders/program_hash_test_utils.py:7-9).
"""
from stark_perpetual_b200._lib import get_context, ints_to_limbs, limbs_to_ints


def compute_hash_chain(data, hash_func=None):
    assert len(data) >= 1, "len(data) for hash chain computation must be >= 1."
    if hash_func is not None:               # an injected hash function: plain fold, as published
        import functools
        return functools.reduce(lambda x, y: hash_func(y, x), data[::-1])
    out, st = get_context().hash_chain_rfold(ints_to_limbs(data), len(data))
    assert st[0] != 1, "hash chain element out of range"
    assert st[0] != 2, "Unhashable input."
    return limbs_to_ints(out)[0]
