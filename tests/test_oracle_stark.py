"""CPU: the prover oracle (oracle/stark.py) on its own -- completeness, the canonical-unpacking soundness fix
(a witness that walks the bits of x + p must be rejected), the query-count floor of verify()."""
import random

import pytest

from oracle import stark
from oracle.params import FIELD_PRIME as P
from oracle.pedersen import pedersen_hash


def _inputs(seed, log_n=9):
    rng = random.Random(seed)
    inst = (1 << log_n) // 512
    return [rng.randrange(1 << 250) for _ in range(5)], [[rng.randrange(1 << 251) for _ in range(inst)] for _ in range(5)]


@pytest.fixture(scope="module")
def honest():
    x0, ys = _inputs(5)
    cols, outs = stark.gen_trace(9, 0, x0, ys)
    return x0, ys, cols, outs, stark.prove_trace(9, 0, x0, outs, cols, n_queries=30)


def test_honest_proof_verifies_and_outputs_are_reference_hashes(honest):
    x0, ys, _cols, outs, proof = honest
    st = stark.verify(proof)
    assert st["outs"] == outs == [pedersen_hash(x0[l], ys[l][0]) for l in range(5)]


def test_noncanonical_unpacking_is_rejected():
    """x and x + p are the same field element; only x is what signature.py:307-317 hashes.  The 251-bit unpacking
    (c6 from row 251 on) leaves the cheating prover no second decomposition."""
    x0, ys = _inputs(6)
    for key, w in (((0, 0, 0), x0[0] + P), ((3, 0, 1), ys[3][0] + P)):
        cols, outs = stark.gen_trace(9, 0, x0, ys, unpack_override={key: w})
        l = key[0]
        assert outs[l] != pedersen_hash(x0[l], ys[l][0])          # the forged statement is false ...
        try:
            proof = stark.prove_trace(9, 0, x0, outs, cols, n_queries=30)
        except ValueError:
            continue                                               # ... the prover may already notice
        with pytest.raises(stark.ProofError):                      # ... and no verifier accepts it
            stark.verify(proof)


def test_inputs_above_2_251_are_outside_the_air():
    x0, ys = _inputs(7)
    x0[2] = (1 << 251) + 5
    with pytest.raises(AssertionError, match="canonical range"):
        stark.gen_trace(9, 0, x0, ys)


def test_verify_enforces_query_floor(honest):
    x0, _ys, cols, outs, _ = honest
    weak = stark.prove_trace(9, 0, x0, outs, cols, n_queries=2)
    with pytest.raises(stark.ProofError, match="queries"):
        stark.verify(weak)
    assert stark.verify(weak, min_queries=2)["n_queries"] == 2


def test_stage_probe_on_cpu_backend():
    """The sampled-point stage checker the 2^20 GPU parity test uses (tests/stage_probe.py), exercised here on the
    oracle-backed CPU stage backend: it must see every stage and agree with it."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cpu_stage_backend import CpuBackend
    from stage_probe import StageChecker
    from stark_perpetual_b200 import prover
    from stark_perpetual_b200._lib import ints_to_limbs
    log_n, chain_log = 9, 0
    x0, ys = _inputs(11)
    cols, outs = stark.gen_trace(log_n, chain_log, x0, ys)
    tr = ints_to_limbs([v for c in cols for v in c])
    be = CpuBackend()
    chk = StageChecker(be, tr, log_n, chain_log, x0, outs, n_groups=24)
    block = be.upload(tr.reshape(25, 1 << log_n, 4))
    proof = prover.prove_sharded(be, prover.TorchComm(0, 1), block, log_n, chain_log, x0, outs, 30, probe=chk)
    chk.assert_complete()
    assert proof == stark.prove_trace(log_n, chain_log, x0, outs, cols, n_queries=30)
    # the checker must notice a wrong value: corrupt the sample's expectation and replay one stage
    pt = chk.points[0]
    chk.h_vals[pt][0] ^= 1
    with pytest.raises(AssertionError):
        chk.on_deep_quotient(gamma=5, layer0=be.felts(8, 1 << log_n))
