"""CPU: the CUDA NTT / LDE tile code (csrc/ntt.cuh) compiled with g++ and run thread-by-thread on the host
against textbook transforms -- catches index / twiddle / planner bugs without a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
BUILD = os.path.join(ROOT, "tests", "host_emul", "build")


# tile shapes (log2 workspace, log2 elements per thread): the shipped kernel shape first, then the radix-8 shape
SHAPES = [(10, 2), (11, 3), (11, 2)]


def _build(name, shape=None):
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(ROOT, "tests", "host_emul", name + ".cpp")
    exe = os.path.join(BUILD, name + ("_%d_%d" % shape if shape else ""))
    defs = ["-DEMUL_LOG_WS=%d" % shape[0], "-DEMUL_LOG_EPT=%d" % shape[1]] if shape else []
    csrc = os.path.join(ROOT, "stark_perpetual_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".inc", ".h"))]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        # SPG_EMUL_LAZY: the host arithmetic reproduces the device's lazy representatives and aborts on any
        # violated bound (fp.cuh), so the butterflies' bound bookkeeping is checked here, without a GPU
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-DSPG_EMUL_LAZY"] + defs + ["-o", exe, src])
    return exe


@pytest.mark.parametrize("args", ["5 0 0", "11 1 0", "12 0 1", "13 1 1", "14 0 1 3", "9 0 1 5",
                                  "10 0 0 -1 1", "10 0 1 -1 1", "9 0 1 -1 2", "13 0 0 -1 2", "8 1 1 -1 1", "20 0 1 2 1"])
@pytest.mark.parametrize("shape", SHAPES)
def test_emulated_ntt_passes(args, shape):
    if shape != SHAPES[0] and args.startswith("20"):
        pytest.skip("2^20 emulation only for the shipped shape")
    exe = _build("emul_ntt", shape)
    out = subprocess.run([exe] + args.split(), capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("args", ["3 3", "10 1", "12 3", "12 3 0", "13 2", "11 3"])
@pytest.mark.parametrize("shape", SHAPES)
def test_emulated_lde(args, shape):
    exe = _build("emul_lde", shape)
    out = subprocess.run([exe] + args.split(), capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


def test_emulated_pedersen_vs_reference_golden(golden):
    """The per-thread Pedersen code of the CUDA kernel (csrc/ec.cuh) run on the host against all golden
    vectors generated from the reference's signature.pedersen_hash."""
    exe = _build("emul_pedersen")
    vec = golden["pedersen"]
    inp = "".join("%s %s\n" % (a[2:], b[2:]) for a, b, _o, _t in vec)
    inp += "%x 1\n" % (2**251 + 17 * 2**192 + 1)      # x == p: out of range
    out = subprocess.run([exe], input=inp, capture_output=True, text=True).stdout.split("\n")
    for (a, b, o, _t), line in zip(vec, out):
        h, st = line.split()
        assert st == "0" and int(h, 16) == int(o, 16), (a, b)
    assert out[len(vec)].split()[1] == "1"


def test_emulated_ecdsa_vs_reference_golden(golden):
    """The per-thread ECDSA code of the CUDA kernel (csrc/ecdsa.cuh: mod-n inversion, square root, the three
    mimic_ec_mult_air loops with every assertion) run on the host against all 65 golden verify vectors
    (True / False / raises) and the 30 reference key pairs."""
    exe = _build("emul_ecdsa")
    lines = []
    for msg, r, s, pub, _res, _tag in golden["verify"]:
        if isinstance(pub, str):
            lines.append("%s %s %s %s -" % (msg[2:], r[2:], s[2:], pub[2:]))
        else:
            lines.append("%s %s %s %s %s" % (msg[2:], r[2:], s[2:], pub[0][2:], pub[1][2:]))
    for priv, _pub in golden["keys"]:
        lines.append("K %s" % priv[2:])
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True).stdout.split()
    nv = len(golden["verify"])
    assert len(out) == nv + len(golden["keys"])
    for v, o in zip(golden["verify"], out[:nv]):
        assert int(o) == v[4], v[5]
    for (_priv, pub), o in zip(golden["keys"], out[nv:]):
        assert int(o, 16) == int(pub, 16)


def test_emulated_sign_vs_golden_and_oracle(golden):
    """The per-thread signing code of the CUDA kernel (csrc/sign.cuh: SHA-256 / HMAC / RFC 6979 nonce, k*G, the
    mod-n finish and the retry rules) run on the host: the reference's four JS deterministic-signature KATs
    (signature.spec.js:96-137), the 14 signatures the reference's sign produced (tests/golden), and seeded
    signatures against the oracle."""
    import random
    from oracle import ecdsa as oec
    exe = _build("emul_ecdsa")
    cases = [(int(m, 16), int(d, 16), 0, int(r, 16), int(s, 16)) for m, d, r, s in golden["sign"] + golden["sign_js_kat"]]
    rng = random.Random(77)
    for seed in (0, 1, 255, 256, 2**32 + 5, 2**64 - 2):
        m, d = rng.randrange(2**251), rng.randrange(1, oec.EC_ORDER)
        cases.append((m, d, seed) + oec.sign(m, d, seed if seed else None))
    for m in (0, 1, 2**248 - 1, 2**248, 2**249 + 3, 2**251 - 1):      # bit lengths around the nibble rule (:119-121)
        d = rng.randrange(1, oec.EC_ORDER)
        cases.append((m, d, 0) + oec.sign(m, d))
    lines = ["S %x %x %d" % (m, d, seed) for m, d, seed, _r, _s in cases]
    lines += ["S %x %x 0" % (2**251, 5), "S %x %x 0" % (5, 0), "S %x %x 0" % (5, oec.EC_ORDER)]
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True).stdout.strip().split("\n")
    assert len(out) == len(lines)
    for (m, d, seed, r, s), o in zip(cases, out):
        st, rr, ss = o.split()
        assert st == "0" and int(rr, 16) == r and int(ss, 16) == s, (hex(m), seed)
    assert [o.split()[0] for o in out[len(cases):]] == ["1", "2", "2"]


_MSG_ARGS = {   # kind -> (code, felt arguments, integer arguments) in the order include/spg.h lists
    "transfer": (4, ("asset_id", "asset_id_fee", "receiver_public_key"),
                 ("sender_position_id", "receiver_position_id", "src_fee_position_id", "nonce", "amount", "max_amount_fee",
                  "expiration_timestamp")),
    "conditional_transfer": (5, ("asset_id", "asset_id_fee", "receiver_public_key", "condition"),
                             ("sender_position_id", "receiver_position_id", "src_fee_position_id", "nonce", "amount",
                              "max_amount_fee", "expiration_timestamp")),
    "withdrawal_to_address": (7, ("asset_id_collateral", "eth_address"), ("position_id", "nonce", "amount", "expiration_timestamp")),
    "price": (100, ("asset_pair", "price"), ("oracle_name", "timestamp")),
}


def _msg_line(kind, fields):
    code, fnames, inames = _MSG_ARGS[kind]
    fv = [int(fields[f], 16) if isinstance(fields[f], str) else fields[f] for f in fnames]
    return "%d %s %s" % (code, " ".join("%x" % v for v in fv), " ".join("%d" % fields[i] for i in inames))


def test_emulated_message_packing_vs_reference_vectors(golden):
    """The device packing code of the transfer / conditional transfer / withdrawal / oracle-price messages
    (csrc/messages.cuh) run on the host: its chain elements, hashed with the oracle's Pedersen hash, must give the 32
    message hashes the reference produced (tests/golden) and its own KATs (perpetual_messages_test.py:22-88); every
    bound the reference asserts must be reported."""
    from oracle.pedersen import pedersen_hash
    exe = _build("emul_messages")
    cases = [(k, f, int(w, 16)) for k, f, w in golden["messages"] if k in _MSG_ARGS]
    pre = golden["messages_precomputed"]
    for want, d in pre["transfer"].items():
        cases.append(("transfer", dict(asset_id=d["assetId"], asset_id_fee=d["assetIdFee"], receiver_public_key=d["receiverPublicKey"],
                                       sender_position_id=d["senderPositionId"], receiver_position_id=d["receiverPositionId"],
                                       src_fee_position_id=d["feePositionId"], nonce=d["nonce"], amount=d["amount"],
                                       max_amount_fee=d["maxAmountFee"], expiration_timestamp=d["expirationTimestamp"]), int(want, 16)))
    for want, d in pre["conditional_transfer"].items():
        cases.append(("conditional_transfer",
                      dict(asset_id=d["assetId"], asset_id_fee=d["assetIdFee"], receiver_public_key=d["receiverPublicKey"],
                           condition=d["condition"], sender_position_id=d["senderPositionId"],
                           receiver_position_id=d["receiverPositionId"], src_fee_position_id=d["srcFeePositionId"],
                           nonce=d["nonce"], amount=d["amount"], max_amount_fee=d["maxAmountFee"],
                           expiration_timestamp=d["expirationTimestamp"]), int(want, 16)))
    for want, d in pre["withdrawal_to_address"].items():
        cases.append(("withdrawal_to_address",
                      dict(asset_id_collateral=d["assetIdCollateral"], eth_address=d["ethAddress"], position_id=d["positionId"],
                           nonce=d["nonce"], amount=d["amount"], expiration_timestamp=d["expirationTimestamp"]), int(want, 16)))
    assert len(cases) >= 35
    # out-of-range variants of the first case of each kind: (field, first value outside the reference's bound)
    bad = []
    limits = {"asset_id": 250, "asset_id_fee": 250, "receiver_public_key": 251, "condition": 251, "nonce": 32,
              "expiration_timestamp": 32, "asset_id_collateral": 250, "eth_address": 160, "asset_pair": 128, "price": 120,
              "oracle_name": 40, "timestamp": 32}
    for kind in _MSG_ARGS:
        base = next(f for k, f, _w in cases if k == kind)
        for name, bits in limits.items():
            if name in base:
                bad.append((kind, dict(base, **{name: 2**bits})))
                ok_edge = dict(base, **{name: 2**bits - 1})
                cases.append((kind, ok_edge, None))
    lines = [_msg_line(k, f) for k, f, _w in cases] + [_msg_line(k, f) for k, f in bad]
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True).stdout.strip().split("\n")
    assert len(out) == len(lines)
    for (kind, fields, want), o in zip(cases, out):
        tok = o.split()
        assert tok[0] == "0", (kind, fields)
        el = [int(t, 16) for t in tok[1:]]
        if want is None:
            continue
        h = pedersen_hash(el[0], el[1])
        for e in el[2:]:
            h = pedersen_hash(h, e)
        assert h == want, kind
    for (kind, fields), o in zip(bad, out[len(cases):]):
        assert o.split()[0] == "1", (kind, fields)


def test_emulated_air_point_lazy_bounds_and_value():
    """The AIR kernel's per-point code (csrc/air_point.cuh) on the host with the lazy-bound checks armed: every
    cell, periodic value, public value, alpha power and inverse zerofier drawn from {0, 1, 2, p-2, p-1, random}, so
    each sum and difference meets its worst case; the result must equal the constraint formulas evaluated with
    Python integers (the same thirteen constraints as oracle/stark.py Air.composition)."""
    import random
    P = 2**251 + 17 * 2**192 + 1
    rng = random.Random(2025)
    exe = _build("emul_air")

    def draw(mode):
        if mode == 0:
            return P - 1
        if mode == 1:
            return 0
        return rng.choice([0, 1, 2, P - 2, P - 1, rng.randrange(P), rng.randrange(P)])
    lines, want = [], []
    for case in range(300):
        mode = case if case < 2 else 2
        px, py, sx, sy = (draw(mode) for _ in range(4))
        alpha = [draw(mode) for _ in range(65)]
        lanes = [[draw(mode) for _ in range(10)] for _ in range(5)]
        iz = [draw(mode) for _ in range(8)]
        acc = [0] * 8
        for l, (X, Y, S, M, I, Xn, Yn, Mn, x0, out) in enumerate(lanes):
            al = alpha[13 * l:13 * l + 13]
            bit = M - 2 * Mn
            c = [bit * (bit - 1), bit * (S * (X - px) - (Y - py)), bit * (S * S - X - px - Xn) + (1 - bit) * (Xn - X),
                 bit * (S * (X - Xn) - Y - Yn) + (1 - bit) * (Yn - Y), I * (X - px) - 1, M, Xn - X, Yn - Y, Mn - X,
                 X - sx, Y - sy, M - x0, X - out]
            groups = [(0, 1, 2, 3), (4,), (5,), (6, 7), (8,), (9, 10), (11,), (12,)]
            for g, ks in enumerate(groups):
                acc[g] += sum(al[k] * c[k] for k in ks)
        want.append(sum(a * z for a, z in zip(acc, iz)) % P)
        vals = [px, py, sx, sy] + alpha + [v for lane in lanes for v in lane] + iz
        lines.append(" ".join("%x" % v for v in vals))
    res = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-500:]
    out = res.stdout.split()
    assert len(out) == len(want)
    for k, (o, w) in enumerate(zip(out, want)):
        assert int(o, 16) == w, k


def test_emulated_math_utils_vs_reference_golden():
    """csrc/ecdsa.cuh ec_op_one (ec_add / ec_double / ec_mult in the reference's evaluation order, with its assertions)
    and fp_sqrt_min, run on the host against vectors generated by the reference's math_utils.py
    (tests/golden/math_utils_golden.json: math_utils.py:36-100)."""
    import json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "math_utils_golden.json")))
    exe = _build("emul_ecdsa")
    lines, want = [], []
    for a, b, res in g["ec_add"]:
        lines.append("E 0 %x %x %x %x" % (a[0], a[1], b[0], b[1])); want.append(res)
    for a, res in g["ec_double"]:
        lines.append("E 1 %x %x 0 0" % (a[0], a[1])); want.append(res)
    for m, a, res in g["ec_mult"]:
        lines.append("E 2 %x %x %x 0" % (a[0], a[1], m)); want.append(res)
    for a, qr, root in g["sqrt_mod"]:
        lines.append("Q %x" % a); want.append((qr, root))
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True).stdout.strip().split("\n")
    assert len(out) == len(want)
    for line, w, src in zip(out, want, lines):
        f = line.split()
        if isinstance(w, tuple):
            assert (f[0] == "0") == w[0], src
            if w[0]:
                assert int(f[1], 16) == w[1], src
        elif w == "AssertionError":
            assert f[0] == "1", src
        else:
            assert f[0] == "0" and [int(f[1], 16), int(f[2], 16)] == list(w), src
