"""The ECDSA-builtin AIR (second AIR): CPU tests of the oracle twin (oracle/stark_ecdsa.py) and of the
CUDA per-point constraint code (csrc/ecdsa_air_point.cuh) compiled with g++ and compared with the oracle.

The reference pins the WITNESS: every block of the trace is one `verify` call walked step by step
(signature.py:176-190, :243-260).  The constraint system is this repo's own (parity unpinned)."""
import random
import subprocess

import pytest

from oracle import ecdsa as oe
from oracle import stark, stark_ecdsa as se
from oracle.params import BETA, FIELD_PRIME as P, SHIFT_POINT
from test_host_emul import _build


@pytest.fixture(scope="module")
def sigs2():
    return se.make_signatures(2, 7)


def test_trace_rows_follow_the_reference_verify(sigs2):
    """the partial sums in the trace are the reference algorithm's: lane C's last partial sum minus the shift point has
    x == r (signature.py:257-260), and every lane's final row equals mimic_ec_mult_air's return value"""
    from oracle.params import EC_GEN, MINUS_SHIFT_POINT
    from oracle.curve import ec_add
    cols = se.gen_trace(9, sigs2)
    for b, (z, r, w, key) in enumerate(sigs2):
        last = 256 * b + 255
        zg = oe.mimic_ec_mult_air(z, EC_GEN, MINUS_SHIFT_POINT)
        rq = oe.mimic_ec_mult_air(r, key, SHIFT_POINT)
        assert (cols[se.APX][last], cols[se.APY][last]) == zg
        assert (cols[se.BPX][last], cols[se.BPY][last]) == rq
        nb = (b + 1) % 2                                        # lane C of signature b lives in the next block
        wb = oe.mimic_ec_mult_air(w, ec_add(zg, rq), SHIFT_POINT)
        assert (cols[se.CPX][256 * nb + 255], cols[se.CPY][256 * nb + 255]) == wb
        assert ec_add(wb, MINUS_SHIFT_POINT)[0] == r
        assert cols[se.T2][256 * nb] == r and cols[se.T1][256 * b + 100] == r


def test_invalid_signature_has_no_trace(sigs2):
    z, r, w, key = sigs2[1]
    with pytest.raises(ValueError):
        se.gen_trace(9, [sigs2[0], (z ^ 1, r, w, key)])
    with pytest.raises(ValueError):
        se.gen_trace(9, [sigs2[0], (z, r, w, (key[0], (key[1] + 1) % P))])          # key off the curve
    with pytest.raises(ValueError):
        se.gen_trace(9, [sigs2[0], (0, r, w, key)])                                  # assert 0 < m


def test_oracle_proof_roundtrip_and_tamper(sigs2):
    proof = se.prove(9, sigs2)
    st = stark.verify(proof)
    assert st["air"] == "ecdsa" and st["msgs"] == [s[0] for s in sigs2] and st["keys"] == [s[3][0] for s in sigs2]
    # a constrained cell of every kind: scalar, partial sum, doubled point, slope, inverse, carrier, non-zero witness
    for col, row in ((se.AM, 3), (se.BPX, 300), (se.CQY, 17), (se.BSD, 250), (se.CI, 255), (se.T2, 400), (se.V1, 0), (se.ASA, 255)):
        with pytest.raises((ValueError, stark.ProofError)):
            stark.verify(se.prove(9, sigs2, corrupt=(col, row, 1)))
    # every instance's public cells are bound: the same proof under another statement (message 0, key 1) is rejected
    for off in (24 + 31, 24 + 3 * 32 + 31):
        bad = bytearray(proof)
        bad[off] ^= 1
        with pytest.raises(stark.ProofError):
            stark.verify(bytes(bad))
    with pytest.raises(stark.ProofError):
        stark.verify(se.prove(9, sigs2, n_queries=12))


def test_emulated_ecdsa_air_point_vs_oracle():
    """csrc/ecdsa_air_point.cuh on the host against EcdsaAir.composition_per on random and edge-value cells"""
    rng = random.Random(77)
    exe = _build("emul_ecdsa_air")
    air = object.__new__(se.EcdsaAir)
    names = ["step", "hold", "zero", "first", "last", "thold"]

    def draw(mode):
        if mode == 0:
            return P - 1
        if mode == 1:
            return 0
        return rng.choice([0, 1, 2, P - 2, P - 1, rng.randrange(P), rng.randrange(P), rng.randrange(P)])
    lines, want = [], []
    for case in range(200):
        mode = case if case < 2 else 2
        gx, gy, fm, fk = (draw(mode) for _ in range(4))
        alpha = [draw(mode) for _ in range(se.N_ALPHA)]
        cur, nxt = [draw(mode) for _ in range(25)], [draw(mode) for _ in range(25)]
        iz = [draw(mode) for _ in range(6)]
        want.append(air.composition_per(cur, nxt, (gx, gy, fm, fk), dict(zip(names, iz)), alpha))
        vals = [gx, gy, SHIFT_POINT[0], SHIFT_POINT[1], BETA, fm, fk] + alpha + cur + nxt + iz
        lines.append(" ".join("%x" % v for v in vals))
    res = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-500:]
    out = res.stdout.split()
    assert len(out) == len(want)
    for k, (o, w) in enumerate(zip(out, want)):
        assert int(o, 16) == w, k


def test_emulated_witness_generator_vs_oracle_trace(sigs2):
    """csrc/ecdsa_air_witness.cuh (Jacobian walks + batched-inversion finish: the three witness kernels' per-thread code) run
    thread by thread on the host == the oracle twin's trace, cell for cell; an invalid signature, an off-curve key and an
    out-of-range scalar raise the status bits the C-ABI turns into errors"""
    exe = _build("emul_ecdsa_air_witness")

    def run(sigs):
        inp = "".join("%x %x %x %x %x\n" % (z, r, w, key[0], key[1]) for z, r, w, key in sigs)
        res = subprocess.run([exe, "9"], input=inp, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr[-500:]
        lines = res.stdout.split()
        return int(lines[0]), [int(v, 16) for v in lines[1:]]
    st, flat = run(sigs2)
    want = se.gen_trace(9, sigs2)
    assert st == 0 and len(flat) == 25 * 512
    for c in range(25):
        assert flat[512 * c:512 * (c + 1)] == want[c], "column %d" % c
    z, r, w, key = sigs2[1]
    assert run([sigs2[0], (z ^ 1, r, w, key)])[0] & 4                                 # verify() == False
    assert run([sigs2[0], (z, r, w, (key[0], (key[1] + 1) % P))])[0] & 1              # key off the curve
    assert run([sigs2[0], (z, r + (1 << 251), w, key)])[0] & 1                        # r outside [1, 2^251)
