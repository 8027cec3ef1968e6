"""Locate and import the REFERENCE's own Python for this path (TEST INFRASTRUCTURE: golden-vector generation, the
`cpu_baseline.kind = "reference"` leg of bench.py, and the "reference tests run unchanged" GPU test).

Where the reference comes from:
  * /root/reference/src            -- the build container (read-only checkout);
  * oracle/_ref/src                -- a staged copy of exactly the files this path touches, made by oracle/stage_ref.py
                                      when /root/reference is present (git-ignored, travels to the GPU box with gpurun).
Nothing under stark_perpetual_b200/ imports this module.

Three harness-side shims are needed because the image lacks `ecdsa`, `web3`, `mypy_extensions` and ships a newer sympy
(SURVEY.md section 8c): `sympy.core.numbers.igcdex` is aliased from sympy.core.intfunc; `ecdsa.rfc6979.generate_k` is the
oracle's RFC 6979 (pinned by the 4 JS KATs of signature.spec.js:96-137 -- signatures "made by the reference" in the golden
files are therefore circular in the nonce, not in anything else); web3 / mypy_extensions are import-only stubs.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CANDIDATES = ["/root/reference/src", os.path.join(_HERE, "_ref", "src")]


def ref_src():
    for c in CANDIDATES:
        if os.path.exists(os.path.join(c, "starkware", "crypto", "signature", "signature.py")):
            return c
    return None


def install_shims():
    import sympy.core.intfunc
    import sympy.core.numbers
    sympy.core.numbers.igcdex = sympy.core.intfunc.igcdex
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import ecdsa as o_ecdsa
    m = types.ModuleType("ecdsa")
    m.rfc6979 = types.ModuleType("ecdsa.rfc6979")
    m.rfc6979.generate_k = o_ecdsa.rfc6979_generate_k
    sys.modules["ecdsa"] = m
    sys.modules["ecdsa.rfc6979"] = m.rfc6979
    me = types.ModuleType("mypy_extensions")
    me.VarArg = lambda t: t
    sys.modules["mypy_extensions"] = me
    w3 = types.ModuleType("web3")
    w3.Web3 = object
    sys.modules["web3"] = w3


def import_reference():
    """Returns (signature, math_utils, perpetual_messages) -- the reference's modules, unmodified.  Raises ImportError when
    neither /root/reference nor the staged copy exists.  Must not be mixed with the compat tree in one process (both
    use the top-level package names `starkware` and `services`)."""
    src = ref_src()
    if src is None:
        raise ImportError("reference sources not found (neither /root/reference/src nor oracle/_ref/src)")
    for name in list(sys.modules):
        if name == "starkware" or name.startswith("starkware.") or name == "services" or name.startswith("services."):
            if not (getattr(sys.modules[name], "__file__", None) or "").startswith(src):
                raise ImportError("module %s already imported from elsewhere (compat tree?)" % name)
    install_shims()
    if src not in sys.path:
        sys.path.insert(0, src)
    from services.perpetual.public import perpetual_messages
    from starkware.crypto.signature import math_utils, signature
    assert signature.__file__.startswith(src)
    return signature, math_utils, perpetual_messages
