"""GPU parity of the ECDSA-builtin AIR (csrc/air_ecdsa.cu, spg_prove_ecdsa) against its CPU twin oracle/stark_ecdsa.py:
witness, composition values and whole proofs bit-exact at 2^9 / 2^10; larger batches accepted by the oracle verifier;
invalid signatures and tampered traces refused.  What the reference pins is the witness (signature.py:176-190, :243-260);
the constraint system is the repo's own (parity unpinned)."""
import numpy as np
import pytest

from oracle import ntt as ontt
from oracle import stark, stark_ecdsa as se
from oracle.params import EC_ORDER, FIELD_PRIME as P

pytestmark = pytest.mark.gpu


def _ctx():
    from stark_perpetual_b200._lib import get_context
    return get_context()


def _limbs(vals):
    from stark_perpetual_b200._lib import ints_to_limbs
    return ints_to_limbs(vals)


def _ints(a):
    from stark_perpetual_b200._lib import limbs_to_ints
    return limbs_to_ints(a)


def _inputs(sigs):
    return (_limbs([s[0] for s in sigs]), _limbs([s[1] for s in sigs]), _limbs([s[2] for s in sigs]),
            _limbs([s[3][0] for s in sigs]), _limbs([s[3][1] for s in sigs]))


@pytest.mark.parametrize("log_n", [9, 10])
def test_trace_composition_and_proof_bit_exact(log_n):
    ctx = _ctx()
    nb = (1 << log_n) >> 8
    sigs = se.make_signatures(nb, 100 + log_n)
    want_cols = se.gen_trace(log_n, sigs)
    trace = ctx.ecdsa_air_trace(log_n, *_inputs(sigs))
    got = _ints(trace)
    n = 1 << log_n
    for c in range(25):
        assert got[c * n:(c + 1) * n] == want_cols[c], "column %d" % c
    # composition polynomial on the cosets 0, 2, 4, 6 with a fixed alpha
    pub = se.public_of(sigs)
    alpha = 0x1234567 * 2**200 + 99
    air = se.EcdsaAir(log_n, pub)
    apows = [pow(alpha, k, P) for k in range(se.N_ALPHA)]
    lde_cols = [ontt.lde(c, stark.LOG_BLOWUP, stark.GEN) for c in want_cols]
    per = air.periodic_lde()
    ins = _inputs(sigs)
    cp = _ints(ctx.air_eval_ecdsa(trace, log_n, ins[0], ins[3], alpha))
    step = 1 if log_n == 9 else 7
    for jj in range(4):
        j = 2 * jj
        for i in range(0, n, step):
            cur = [lde_cols[c][j][i] for c in range(25)]
            nxt = [lde_cols[c][j][(i + 1) % n] for c in range(25)]
            want = air.composition_per(cur, nxt, per(j, i), air.inv_zerofiers(stark.lde_point(log_n, j, i)), apows)
            assert cp[jj * n + i] == want, (jj, i)
    # whole proof
    proof = ctx.prove_ecdsa(trace, log_n, ins[0], ins[3], 30)
    assert proof == stark.prove_air(air, want_cols, 30)
    st = stark.verify(proof)
    assert st["air"] == "ecdsa" and st["msgs"] == [s[0] for s in sigs] and st["keys"] == [s[3][0] for s in sigs]


def _gpu_signatures(ctx, count, seed):
    """count signatures made on the device (spg_sign_batch), keys as points"""
    rng = np.random.default_rng(seed)
    import random
    r_ = random.Random(seed)
    privs = [r_.randrange(1, 1 << 250) for _ in range(count)]
    msgs = [r_.randrange(1, 1 << 251) for _ in range(count)]
    kx, ky, st = ctx.private_to_stark_key(_limbs(privs), want_y=True)
    assert not st.any()
    r, s, st = ctx.sign(_limbs(msgs), _limbs(privs))
    assert not st.any()
    del rng
    return msgs, _ints(r), _ints(s), list(zip(_ints(kx), _ints(ky)))


@pytest.mark.parametrize("log_n", [14, 17, 20])
def test_larger_batches_verify_under_the_oracle(log_n):
    from stark_perpetual_b200.ecdsa_air import prove_signatures
    ctx = _ctx()
    count = (1 << log_n) >> 8
    msgs, r, s, keys = _gpu_signatures(ctx, count, 5 + log_n)
    proof, ln = prove_signatures(msgs, r, s, keys, ctx=ctx)
    assert ln == log_n
    st = stark.verify(proof)
    assert st["log_n"] == log_n and st["msgs"] == msgs and st["keys"] == [k[0] for k in keys]
    # the witness rows follow the reference's verify: spot-check one signature against the oracle's own verify
    from oracle import ecdsa as oe
    assert oe.verify(msgs[3], r[3], s[3], keys[3])


def test_invalid_inputs_are_refused():
    from stark_perpetual_b200._lib import SpgError
    ctx = _ctx()
    sigs = se.make_signatures(2, 9)
    z, r, w, key = sigs[1]
    for bad in ((z ^ 1, r, w, key),                           # verify() == False
                (z, r, w, (key[0], (key[1] + 1) % P)),        # key off the curve
                (0, r, w, key),                               # assert 0 < m (signature.py:180)
                (z, r + (1 << 251), w, key)):                 # r outside [1, 2^251) (signature.py:221)
        with pytest.raises(SpgError):
            ctx.ecdsa_air_trace(9, *_inputs([sigs[0], bad]))
    # a tampered trace has no proof (the prover checks the composition at the out-of-domain point)
    ins = _inputs(sigs)
    trace = ctx.ecdsa_air_trace(9, *ins)
    for col, row in ((se.BM, 0), (se.CPX, 300), (se.T2, 260), (se.CSA, 511)):
        t2 = trace.copy()
        t2[col * 512 + row, 0] ^= np.uint64(1)
        with pytest.raises(SpgError):
            ctx.prove_ecdsa(t2, 9, ins[0], ins[3], 30)
    # a public input that is not what the trace holds (message of signature 1, key of signature 0)
    for which in (0, 3):
        pub = [ins[0].copy(), ins[3].copy()]
        pub[which == 3][1 if which == 0 else 0, 0] ^= np.uint64(1)
        with pytest.raises(SpgError):
            ctx.prove_ecdsa(trace, 9, pub[0], pub[1], 30)
    assert EC_ORDER > 0


def test_file_level_prover_cli_ecdsa(tmp_path):
    """the prover CLI (flag set of the one that follows cairo-run, cairo_cmake_rules.cmake:72-110) over the second AIR:
    instances in the public input file, signatures in the private input file, proof file out"""
    import json
    import os
    import subprocess
    import sys
    sigs = se.make_signatures(4, 31)
    from oracle import ecdsa as oe
    inst = [{"msg": hex(z), "key": hex(key[0])} for z, _r, _w, key in sigs]
    prv = [{"r": hex(r), "s": hex(pow(w, -1, EC_ORDER)), "key_y": hex(key[1])} for _z, r, w, key in sigs]
    assert oe.verify(sigs[0][0], sigs[0][1], int(prv[0]["s"], 16), sigs[0][3])
    (tmp_path / "public.json").write_text(json.dumps({"air": "ecdsa", "instances": inst}))
    (tmp_path / "private.json").write_text(json.dumps({"signatures": prv}))
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    cmd = [sys.executable, "-m", "stark_perpetual_b200.cpu_air_prover", "--out_file", str(tmp_path / "proof.bin"),
           "--private_input_file", str(tmp_path / "private.json"), "--public_input_file", str(tmp_path / "public.json")]
    out = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    st = stark.verify((tmp_path / "proof.bin").read_bytes())
    assert st["msgs"] == [s[0] for s in sigs] and st["keys"] == [s[3][0] for s in sigs] and st["log_n"] == 10
    # a signature that does not verify: no proof, a message instead
    prv[2]["r"] = hex(int(prv[2]["r"], 16) ^ 4)
    (tmp_path / "private.json").write_text(json.dumps({"signatures": prv}))
    out = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode != 0 and "does not verify" in out.stderr


@pytest.mark.parametrize("log_n,n_points", [(12, 256), (20, 1024)])
def test_composition_at_sampled_points_headline_size(log_n, n_points):
    """the constraint kernel at the bench's trace size (4096 signatures): composition values at sampled points of the four
    even cosets against the twin's constraint code fed with the C oracle's LDE of the same trace, the twin's own public
    and periodic columns and zerofiers"""
    import random
    from oracle import clib
    ctx = _ctx()
    n = 1 << log_n
    count = n >> 8
    msgs, r, s, keys = _gpu_signatures(ctx, count, 77)
    from stark_perpetual_b200.ecdsa_air import air_inputs
    m_, r_, w_, kx_, ky_ = air_inputs(msgs, r, s, keys)
    trace = ctx.ecdsa_air_trace(log_n, m_, r_, w_, kx_, ky_)
    alpha = 0x3039 * 2**190 + 12345
    cp = ctx.air_eval_ecdsa(trace, log_n, m_, kx_, alpha)                      # (4 N, 4)
    rng = random.Random(5)
    # the block-boundary rows carry most constraints' interesting cases: force some of them into the sample
    rows = [rng.randrange(n) for _ in range(n_points - 64)] + [256 * rng.randrange(count) + t for t in (0, 250, 251, 255) for _ in range(16)]
    pts = [(2 * rng.randrange(4), i) for i in rows]
    want_pts = sorted({(j, i) for j, i in pts} | {(j, (i + 1) % n) for j, i in pts})
    jj = np.array([p[0] for p in want_pts]); ii = np.array([p[1] for p in want_pts])
    clib.use_all_cores()
    t_vals = {pt: [0] * 25 for pt in want_pts}
    tr = trace.reshape(25, n, 4)
    for c0 in range(0, 25, 13):
        step = min(13, 25 - c0)
        lde = clib.lde(np.ascontiguousarray(tr[c0:c0 + step]).reshape(-1, 4), log_n, step, 3).reshape(8, step, n, 4)
        for c in range(step):
            for pt, v in zip(want_pts, _ints(np.ascontiguousarray(lde[jj, c, ii]))):
                t_vals[pt][c0 + c] = v
        del lde
    air = se.EcdsaAir(log_n, list(zip(msgs, [k[0] for k in keys])))
    apows = [pow(alpha, k, P) for k in range(se.N_ALPHA)]
    got = _ints(np.ascontiguousarray(cp.reshape(4, n, 4)[np.array([p[0] // 2 for p in pts]), np.array([p[1] for p in pts])]))
    for (j, i), g in zip(pts, got):
        x = stark.lde_point(log_n, j, i)
        want = air.composition_per(t_vals[(j, i)], t_vals[(j, (i + 1) % n)], air.periodic_at(x), air.inv_zerofiers(x), apows)
        assert g == want, (j, i)


def test_sharded_prover_world1_equals_monolithic():
    """spg_prove_ecdsa_sharded with a one-rank communicator (no NCCL needed) == spg_prove_ecdsa; the 2 / 4 / 8-GPU form of the
    same check is tools/multi_gpu_check.py (profiles/*multi_gpu_check*)"""
    from stark_perpetual_b200._lib import Context
    ctx = Context(0)
    sigs = se.make_signatures(4, 41)
    ins = _inputs(sigs)
    trace = ctx.ecdsa_air_trace(10, *ins)
    ctx.comm_init(0, 1)
    assert ctx.prove_ecdsa_sharded(trace, 10, ins[0], ins[3], 30) == ctx.prove_ecdsa(trace, 10, ins[0], ins[3], 30)
    ctx.close()
