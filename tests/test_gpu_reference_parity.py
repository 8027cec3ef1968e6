"""GPU parity against REFERENCE-generated vectors and the reference's OWN tests:

  * compat math_utils (spg_ec_op_batch, spg_field_sqrt_batch) vs tests/golden/math_utils_golden.json
    (math_utils.py:36-100, incl. every assertion case);
  * BASELINE.json configs[4] / SURVEY.md section 8(d) cfg-5: the 512-order stratified sample of
    tests/golden/orders_golden.json -- message hashes and verify() outcomes made by the reference;
  * the reference's test files for this path (perpetual_messages_test.py:22-88, stark_cli_test.py:43-146) executed
    UNCHANGED with the GPU-backed compat tree ahead of the reference on PYTHONPATH.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import refenv
from stark_perpetual_b200._lib import ints_to_limbs, limbs_to_ints

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
COMPAT = os.path.join(ROOT, "stark_perpetual_b200", "compat")
P = 2**251 + 17 * 2**192 + 1


@pytest.fixture(scope="module")
def compat_modules():
    sys.path.insert(0, COMPAT)
    try:
        from services.perpetual.public import perpetual_messages as pm
        from starkware.crypto.signature import math_utils as mu
        assert mu.__file__.startswith(COMPAT) and pm.__file__.startswith(COMPAT)
        yield mu, pm
    finally:
        sys.path.remove(COMPAT)


def test_math_utils_vs_reference_vectors(ctx, compat_modules):
    mu, _pm = compat_modules
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "math_utils_golden.json")))

    def check(call, want):
        if want == "AssertionError":
            with pytest.raises(AssertionError):
                call()
        else:
            assert list(call()) == list(want)
    for a, b, res in g["ec_add"]:
        check(lambda: mu.ec_add(tuple(a), tuple(b), P), res)
    for a, res in g["ec_double"]:
        check(lambda: mu.ec_double(tuple(a), 1, P), res)
    for m, a, res in g["ec_mult"]:
        check(lambda: mu.ec_mult(m, tuple(a), 1, P), res)
    for n, m, res in g["div_mod"]:
        assert mu.div_mod(n, m, P) == res
    for a, qr, root in g["sqrt_mod"]:
        assert mu.is_quad_residue(a, P) == qr
        if qr:
            assert mu.sqrt_mod(a, P) == root
    # batched forms: statuses per element through the C-ABI
    good = [(tuple(a), tuple(b), tuple(r)) for a, b, r in g["ec_add"] if r != "AssertionError"]
    assert mu.ec_add_batch([x[0] for x in good], [x[1] for x in good]) == [x[2] for x in good]
    xy = ints_to_limbs([c for a, _b, _r in g["ec_add"] for c in a]).reshape(-1, 8)
    xy2 = ints_to_limbs([c for _a, b, _r in g["ec_add"] for c in b]).reshape(-1, 8)
    _out, st = ctx.ec_op(0, xy, xy2)
    assert st.tolist() == [1 if r == "AssertionError" else 0 for _a, _b, r in g["ec_add"]]
    ms = [m for m, _a, _r in g["ec_mult"]] + [0]
    pts = [a for _m, a, _r in g["ec_mult"]] + [g["ec_mult"][0][1]]
    _out, st = ctx.ec_op(2, ints_to_limbs([c for a in pts for c in a]).reshape(-1, 8), ints_to_limbs(ms))
    assert st.tolist() == [1 if r == "AssertionError" else 0 for _m, _a, r in g["ec_mult"]] + [3]
    # a coordinate that is not a field element
    _out, st = ctx.ec_op(1, ints_to_limbs([P, 5]).reshape(-1, 8))
    assert st.tolist() == [2]
    # a larger batch (several scratch chunks) stays consistent with the generator-table kernel
    rng = np.random.default_rng(5)
    ks = [int(x) for x in rng.integers(1, 2**62, size=9000)]
    gen = g["ec_mult"][0][1]
    out, st = ctx.ec_op(2, np.tile(ints_to_limbs(gen).reshape(1, 8), (len(ks), 1)), ints_to_limbs(ks))
    assert not st.any()
    if gen == list(map(int, gen)) and False:
        pass
    xs = limbs_to_ints(out.reshape(-1, 4))[0::2]
    want = mu.ec_mult_batch(ks[:3] + ks[-3:], [tuple(gen)] * 6)
    assert [xs[0], xs[1], xs[2], xs[-3], xs[-2], xs[-1]] == [w[0] for w in want]


def test_limit_orders_512_reference_sample(ctx, compat_modules):
    """cfg-5 stratified sample: 256 validly signed + 256 corrupted orders; hashes and verify outcomes by the reference."""
    _mu, pm = compat_modules
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "orders_golden.json")))["cases"]
    assert len(cases) == 512
    orders = [{k: int(v, 16) for k, v in c["order"].items()} for c in cases]
    msgs = pm.get_limit_order_msg_batch(orders)
    assert [hex(m) for m in msgs] == [c["msg"] for c in cases]
    rs, ss, keys = ([int(c[k], 16) for c in cases] for k in ("r", "s", "pub"))
    st = ctx.limit_order_verify(pm._order_arrays(orders), ints_to_limbs(rs), ints_to_limbs(ss), ints_to_limbs(keys))
    want = [c["verify"] for c in cases]
    bad = [(i, cases[i]["kind"], cases[i]["what"], int(st[i]), want[i]) for i in range(512) if int(st[i]) != want[i]]
    assert not bad, bad[:5]
    assert sum(1 for w in want if w == 1) == 256 and any(w == 2 for w in want)
    # the plain ECDSA entry point on the reference's message hashes gives the same answers
    st2 = ctx.ecdsa_verify(ints_to_limbs(msgs), ints_to_limbs(rs), ints_to_limbs(ss), ints_to_limbs(keys), None)
    assert st2.tolist() == want


@pytest.mark.parametrize("test_file", ["perpetual_messages_test.py", "stark_cli_test.py"])
def test_reference_tests_run_unchanged(ctx, test_file):
    """The reference's own test module, byte for byte, with `starkware.*` / `services.*` resolved to the compat tree
    (libspg underneath): PYTHONPATH = repo : compat : <reference src>."""
    src = refenv.ref_src()
    if src is None:
        pytest.skip("reference sources not staged (oracle/stage_ref.py needs /root/reference)")
    test_path = os.path.join(src, "services", "perpetual", "public", test_file)
    if test_file == "stark_cli_test.py" and not os.path.exists(os.path.join(os.path.dirname(test_path), "signature_test_data.json")):
        # in the read-only checkout the fixtures are not beside the test (Bazel copies them); use the staged tree
        staged = os.path.join(ROOT, "oracle", "_ref", "src")
        if not os.path.exists(os.path.join(staged, "services", "perpetual", "public", "signature_test_data.json")):
            pytest.skip("stark_cli_test.py needs its fixtures beside it: run oracle/stage_ref.py")
        src = staged
        test_path = os.path.join(src, "services", "perpetual", "public", test_file)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, COMPAT, src]))
    # which modules does that path resolve to?  (must be the compat tree, not the reference's own implementation)
    probe = subprocess.run([sys.executable, "-c",
                            "import services.perpetual.public.perpetual_messages as a, starkware.crypto.signature.signature as b;"
                            "print(a.__file__); print(b.__file__)"], env=env, capture_output=True, text=True, cwd=ROOT)
    assert probe.returncode == 0, probe.stderr
    assert all(line.startswith(COMPAT) for line in probe.stdout.split()), probe.stdout
    out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--import-mode=importlib",
                          "--rootdir", os.path.dirname(test_path), "-c", "/dev/null", test_path],
                         env=env, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert " passed" in out.stdout and "failed" not in out.stdout, out.stdout[-2000:]


def test_reference_program_hash_test_runs_unchanged(ctx, tmp_path):
    """Row f-2: the reference's program_hash_test.py + program_hash_test_utils.py (:7-21), byte for byte, against compat's
    compute_program_hash_chain (libspg underneath).  The compiled perpetual program cannot be produced here (no Cairo
    compiler), so the test directory gets a SYNTHETIC compiled-program JSON and the hash the oracle restatement gives for it;
    a wrong expected hash must make the reference's assertion fire."""
    import random
    import shutil
    from oracle import hash_chain as ohc
    src = refenv.ref_src()
    if src is None or not os.path.exists(os.path.join(src, "services", "perpetual", "cairo", "program_hash_test.py")):
        staged = os.path.join(ROOT, "oracle", "_ref", "src")
        if not os.path.exists(os.path.join(staged, "services", "perpetual", "cairo", "program_hash_test.py")):
            pytest.skip("reference program-hash test not staged")
        src = staged
    rng = random.Random(99)
    builtins = ["output", "pedersen", "range_check", "ecdsa", "bitwise"]
    data = [rng.randrange(P) for _ in range(300)]
    work = tmp_path / "services" / "perpetual" / "cairo"
    work.mkdir(parents=True)
    shutil.copyfile(os.path.join(src, "services", "perpetual", "cairo", "program_hash_test.py"), work / "program_hash_test.py")
    (work / "perpetual_cairo_compiled.json").write_text(json.dumps(
        {"builtins": builtins, "data": [hex(v) for v in data], "identifiers": {"__main__.main": {"pc": 42, "type": "function"}},
         "prime": hex(P)}))
    want = ohc.compute_program_hash_chain(builtins, 42, data)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, COMPAT, src]))
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--import-mode=importlib", "--rootdir", str(work),
           "-c", "/dev/null", str(work / "program_hash_test.py")]
    for expected, ok in ((hex(want), True), (hex(want ^ 1), False)):
        (work / "program_hash.json").write_text(json.dumps({"program_hash": expected}, indent=4) + "\n")
        out = subprocess.run(cmd, env=env, capture_output=True, text=True, cwd=ROOT, timeout=600)
        assert (out.returncode == 0) == ok, out.stdout[-2000:] + out.stderr[-1000:]
        if not ok:
            assert "Wrong program hash" in out.stdout
