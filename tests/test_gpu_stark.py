"""GPU parity: every prover stage and the whole proof against oracle/stark.py (bit-exact), plus acceptance of
larger proofs by the oracle verifier and rejection of tampered traces."""
import hashlib
import random

import numpy as np
import pytest

from conftest import rand_felts
from oracle import stark
from oracle.params import FIELD_PRIME as P, R_MOD_P
from stark_perpetual_b200 import SpgError
from stark_perpetual_b200._lib import ints_to_limbs, limbs_to_ints

pytestmark = pytest.mark.gpu


def make_inputs(log_n, seed):
    rng = random.Random(seed)
    inst = (1 << log_n) // 512
    x0 = [rng.randrange(P) for _ in range(5)]
    ys = [[rng.randrange(P) for _ in range(inst)] for _ in range(5)]
    return x0, ys


def ys_limbs(ys):
    return ints_to_limbs([v for lane in ys for v in lane])


@pytest.mark.parametrize("n_cols,rows", [(1, 8), (1, 64), (4, 16), (25, 32), (3, 1024)])
def test_merkle_commit_vs_oracle(ctx, n_cols, rows):
    tab = rand_felts(8 * n_cols * rows, 900 + rows)
    vals = limbs_to_ints(tab)
    # oracle hashes ser(v) = to_bytes(v * R); feed it v / R so that the hashed bytes are our raw values
    rinv = pow(R_MOD_P, -1, P)
    table = [[[vals[(j * n_cols + c) * rows + i] * rinv % P for i in range(rows)] for c in range(n_cols)] for j in range(8)]
    levels = stark.merkle_levels([stark.H(x) for x in stark.table_leaves(table, rows)])
    root, tree = ctx.merkle_commit(tab, n_cols, rows, want_tree=True)
    assert root == levels[-1][0]
    flat = [h for lvl in levels for h in lvl]
    assert [bytes(t) for t in tree] == flat


@pytest.mark.parametrize("log_n,chain_log", [(9, 0), (10, 1), (11, 0)])
def test_trace_generation_vs_oracle(ctx, log_n, chain_log):
    x0, ys = make_inputs(log_n, 40 + log_n)
    cols, outs = stark.gen_trace(log_n, chain_log, x0, ys)
    tr = ctx.pedersen_chain_trace(log_n, chain_log, x0, ys_limbs(ys))
    assert limbs_to_ints(tr) == [v for c in cols for v in c]


def test_trace_generation_rejects_out_of_range(ctx):
    x0, ys = make_inputs(9, 3)
    ys[2][0] = P
    with pytest.raises(SpgError):
        ctx.pedersen_chain_trace(9, 0, x0, ys_limbs(ys))


@pytest.mark.parametrize("log_n,chain_log", [(9, 0), (10, 1)])
def test_air_composition_vs_oracle(ctx, log_n, chain_log):
    x0, ys = make_inputs(log_n, 50 + log_n)
    cols, outs = stark.gen_trace(log_n, chain_log, x0, ys)
    dbg = {}
    proof = stark.prove_trace(log_n, chain_log, x0, outs, cols, n_queries=3, debug=dbg)
    # alpha is the first challenge after the trace root; recover it by replaying the channel
    ch = stark.Channel(stark.public_seed(log_n, chain_log, 3, x0, outs))
    off = 4 + 5 * 4 + 10 * 32
    ch.absorb(proof[off:off + 32])
    alpha = ch.draw_felt()
    n = 1 << log_n
    tr = ints_to_limbs([v for c in cols for v in c])
    cp = limbs_to_ints(ctx.air_eval(tr, log_n, chain_log, x0, outs, alpha))
    want = dbg["cp"]
    for jj in range(4):
        assert cp[jj * n:(jj + 1) * n] == [want[jj + 4 * i] for i in range(n)], jj


@pytest.mark.parametrize("log_n,chain_log,nq", [(9, 0, 5), (10, 1, 7)])
def test_proof_bit_exact_vs_oracle_prover(ctx, log_n, chain_log, nq):
    x0, ys = make_inputs(log_n, 60 + log_n)
    cols, outs = stark.gen_trace(log_n, chain_log, x0, ys)
    want = stark.prove_trace(log_n, chain_log, x0, outs, cols, n_queries=nq)
    tr = ints_to_limbs([v for c in cols for v in c])
    got = ctx.prove(tr, log_n, chain_log, x0, n_queries=nq)
    assert len(got) == len(want)
    assert got == want
    st = stark.verify(got, min_queries=nq)
    assert st["outs"] == outs and st["x0"] == x0


@pytest.mark.parametrize("log_n,chain_log", [(12, 2), (13, 0), (15, 3), (16, 2)])
def test_larger_proofs_accepted_by_oracle_verifier(ctx, log_n, chain_log):
    x0, ys = make_inputs(log_n, 70 + log_n)
    tr = ctx.pedersen_chain_trace(log_n, chain_log, x0, ys_limbs(ys))
    proof = ctx.prove(tr, log_n, chain_log, x0, n_queries=30)
    st = stark.verify(proof)
    assert st["log_n"] == log_n and st["x0"] == x0
    # the public outputs are the chained Pedersen hashes: recompute lane 0's last segment with the oracle
    from oracle.pedersen import pedersen_hash
    inst = (1 << log_n) // 512
    h = x0[0]
    for q in range(inst - (1 << chain_log), inst):
        h = pedersen_hash(h, ys[0][q])
    assert st["outs"][0] == h
    # tampering with any byte of the proof must be rejected
    rng = random.Random(log_n)
    for _ in range(4):
        bad = bytearray(proof)
        bad[rng.randrange(64, len(bad))] ^= 1 << rng.randrange(8)
        with pytest.raises(stark.ProofError):
            stark.verify(bytes(bad))


def test_invalid_trace_is_refused(ctx):
    log_n, chain_log = 10, 1
    x0, ys = make_inputs(log_n, 81)
    tr = ctx.pedersen_chain_trace(log_n, chain_log, x0, ys_limbs(ys))
    n = 1 << log_n
    for col, row in ((0, 100), (1, 300), (3, 5), (4, 17), (8, 252), (0, 0), (20, n - 1)):
        bad = tr.copy()
        bad[col * n + row, 0] ^= np.uint64(1)
        with pytest.raises(SpgError, match="does not satisfy the AIR"):
            ctx.prove(bad, log_n, chain_log, x0, n_queries=4)
    # wrong public seed
    with pytest.raises(SpgError, match="does not satisfy the AIR"):
        ctx.prove(tr, log_n, chain_log, [x0[0] + 1] + x0[1:], n_queries=4)


def test_noncanonical_unpacking_is_refused(ctx):
    """A witness that walks the bits of x + p (same field element, different integer: signature.py:307 hashes the integer)
    must not yield a proof; inputs >= 2^251 are refused by the witness generator (canonical 251-bit unpacking)."""
    log_n, chain_log = 9, 0
    rng = random.Random(77)
    x0 = [rng.randrange(1 << 250) for _ in range(5)]
    ys = [[rng.randrange(1 << 251)] for _ in range(5)]
    for key, w in (((0, 0, 0), x0[0] + P), ((4, 0, 1), ys[4][0] + P)):
        cols, outs = stark.gen_trace(log_n, chain_log, x0, ys, unpack_override={key: w})
        tr = ints_to_limbs([v for c in cols for v in c])
        with pytest.raises(SpgError, match="does not satisfy the AIR"):
            ctx.prove(tr, log_n, chain_log, x0, n_queries=30)
    cols, outs = stark.gen_trace(log_n, chain_log, x0, ys)
    stark.verify(ctx.prove(ints_to_limbs([v for c in cols for v in c]), log_n, chain_log, x0, n_queries=30))
    big = list(x0)
    big[1] = (1 << 251) + 12345
    with pytest.raises(SpgError, match="2\\^251"):
        ctx.pedersen_chain_trace(log_n, chain_log, big, ys_limbs(ys))


@pytest.mark.parametrize("log_n,chain_log", [(10, 1), (14, 2)])
def test_stage_driver_equals_monolithic_prover(ctx, log_n, chain_log):
    """The multi-GPU driver (prover.prove_sharded over the spg_stage_* entry points) with world = 1 must
    produce exactly the bytes of spg_prove."""
    from stark_perpetual_b200 import prover
    x0, ys = make_inputs(log_n, 90 + log_n)
    tr = ctx.pedersen_chain_trace(log_n, chain_log, x0, ys_limbs(ys))
    want = ctx.prove(tr, log_n, chain_log, x0, n_queries=9)
    pv = prover.Prover(ctx)
    block, outs = pv.shard_host_trace(tr, log_n)
    got = pv.prove_sharded_device(block, log_n, chain_log, x0, outs, n_queries=9)
    assert got == want
    stark.verify(got, min_queries=9)


def test_file_level_prover_cli(ctx, tmp_path):
    """python -m stark_perpetual_b200.cpu_air_prover with the flag set of the prover CLI that follows cairo-run in the
    reference's build (cairo_cmake_rules.cmake:72-110): files in, proof file out, accepted by the oracle verifier."""
    import json
    import os
    import subprocess
    import sys
    log_n, chain_log = 10, 1
    x0, ys = make_inputs(log_n, 123)
    ys_limbs(ys).astype("<u8").tofile(tmp_path / "ys.bin")
    (tmp_path / "public.json").write_text(json.dumps({"log_n": log_n, "chain_log": chain_log, "x0": [hex(v) for v in x0]}))
    (tmp_path / "private.json").write_text(json.dumps({"ys_path": str(tmp_path / "ys.bin")}))
    (tmp_path / "params.json").write_text(json.dumps({"n_queries": 30}))
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    out = subprocess.run([sys.executable, "-m", "stark_perpetual_b200.cpu_air_prover", "--out_file", str(tmp_path / "proof.bin"),
                          "--private_input_file", str(tmp_path / "private.json"), "--public_input_file", str(tmp_path / "public.json"),
                          "--parameter_file", str(tmp_path / "params.json")], cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    proof = (tmp_path / "proof.bin").read_bytes()
    st = stark.verify(proof)
    pub = json.loads((tmp_path / "proof.bin.public.json").read_text())
    assert st["x0"] == x0 and [hex(v) for v in st["outs"]] == pub["outs"]
    # the trace-file form gives the same proof
    tr = ctx.pedersen_chain_trace(log_n, chain_log, x0, ys_limbs(ys))
    tr.astype("<u8").tofile(tmp_path / "trace.bin")
    (tmp_path / "private2.json").write_text(json.dumps({"trace_path": str(tmp_path / "trace.bin")}))
    out = subprocess.run([sys.executable, "-m", "stark_perpetual_b200.cpu_air_prover", "--out_file", str(tmp_path / "proof2.bin"),
                          "--private_input_file", str(tmp_path / "private2.json"), "--public_input_file", str(tmp_path / "public.json")],
                         cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert (tmp_path / "proof2.bin").read_bytes() == proof
