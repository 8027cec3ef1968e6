"""cairo-run artefact ingestion (scaffolding) and builtin-segment checks on the device.

The reference's build rule runs `cairo-run --proof_mode` on the perpetual program and leaves three files for the prover
(src/starkware/cairo/lang/cairo_cmake_rules.cmake:72-84, :94-110):

    <name>_trace.bin          per step 3 x u64 little-endian (ap, fp, pc)
    <name>_memory.bin         per cell  u64 little-endian address || 32-byte little-endian value
    <name>_public_input.json  layout, n_steps, rc_min / rc_max, memory_segments {name: {begin_addr, stop_ptr}},
                              public_memory [{address, value (hex), page}]

The binary formats are cairo-lang's (an un-vendored dependency: scripts/requirements-gen.txt:2) -- EXTERNAL, restated
from its published writer; nothing in the reference tree pins them, so the parser is tested on files this module writes
itself (tests/test_cairo_artifacts.py).  What is done with them here:

  * parse / write the three files (numpy, no per-cell Python loops);
  * slice the BUILTIN SEGMENTS of the perpetual layout (main.cairo:1 `%builtins output pedersen range_check ecdsa bitwise`)
    out of the memory: the Pedersen builtin keeps 3 cells per instance (x, y, result), the ECDSA builtin 2 (public key,
    message) -- cairo-lang's instance layouts, EXTERNAL;
  * recompute every Pedersen instance on the GPU (spg_pedersen_hash2_batch) and compare with the result cells -- the
    consistency check of the witness for the largest builtin of a perpetual batch, and check the
    ECDSA instances against signatures supplied by the private input (spg_ecdsa_verify_batch).

  * prove the ECDSA builtin segment with the repo's ECDSA-builtin AIR (`prove_ecdsa_builtin`): one proof whose public input
    is the segment's (message, key) cells.

The Cairo CPU / memory / range-check AIR itself is NOT implemented (DESIGN.md "Out of scope"): proving these traces needs
the Cairo layout's constraint system, which is cairo-lang / Stone material and absent from the reference.
"""
import json

import numpy as np

from ._lib import FIELD_PRIME, get_context

PEDERSEN_CELLS = 3      # x, y, result            (cairo-lang pedersen builtin instance, EXTERNAL)
ECDSA_CELLS = 2         # public key, message     (cairo-lang ecdsa builtin instance, EXTERNAL)

_TRACE_DT = np.dtype([("ap", "<u8"), ("fp", "<u8"), ("pc", "<u8")])
_MEM_DT = np.dtype([("addr", "<u8"), ("value", "<u8", (4,))])


# ------------------------------------------------------------------ files
def read_trace(path):
    """-> structured array with fields ap, fp, pc (one row per step)"""
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size % _TRACE_DT.itemsize:
        raise ValueError("trace file size %d is not a multiple of %d bytes" % (raw.size, _TRACE_DT.itemsize))
    return raw.view(_TRACE_DT)


def write_trace(path, ap, fp, pc):
    t = np.empty(len(ap), dtype=_TRACE_DT)
    t["ap"], t["fp"], t["pc"] = ap, fp, pc
    t.tofile(path)


class Memory:
    """Sparse Cairo memory: sorted addresses and their 256-bit values (4 x u64 little-endian limbs)."""

    def __init__(self, addr, values):
        order = np.argsort(addr, kind="stable")
        self.addr = np.ascontiguousarray(addr[order], dtype=np.uint64)
        self.values = np.ascontiguousarray(values[order], dtype=np.uint64).reshape(-1, 4)
        if self.addr.size > 1 and (np.diff(self.addr.astype(np.int64)) == 0).any():
            raise ValueError("memory file assigns an address twice")

    def __len__(self):
        return self.addr.size

    def get(self, addresses):
        """values at `addresses` ((k, 4) uint64); KeyError if one is missing"""
        a = np.asarray(addresses, dtype=np.uint64)
        pos = np.searchsorted(self.addr, a)
        ok = (pos < self.addr.size)
        ok[ok] = self.addr[pos[ok]] == a[ok]
        if not ok.all():
            raise KeyError("address %d is not in the memory file" % int(a[~ok][0]))
        return self.values[pos]

    def segment(self, begin, stop):
        """dense view of [begin, stop): (values (stop - begin, 4), present mask)"""
        n = int(stop) - int(begin)
        out = np.zeros((n, 4), dtype=np.uint64)
        lo, hi = np.searchsorted(self.addr, [begin, stop])
        rel = (self.addr[lo:hi] - np.uint64(begin)).astype(np.int64)
        out[rel] = self.values[lo:hi]
        mask = np.zeros(n, dtype=bool)
        mask[rel] = True
        return out, mask


def read_memory(path):
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size % _MEM_DT.itemsize:
        raise ValueError("memory file size %d is not a multiple of %d bytes" % (raw.size, _MEM_DT.itemsize))
    m = raw.view(_MEM_DT)
    return Memory(m["addr"].copy(), m["value"].copy())


def write_memory(path, addr, values):
    m = np.empty(len(addr), dtype=_MEM_DT)
    m["addr"] = addr
    m["value"] = np.asarray(values, dtype=np.uint64).reshape(-1, 4)
    m.tofile(path)


def read_public_input(path):
    with open(path) as f:
        pub = json.load(f)
    for key in ("layout", "n_steps", "memory_segments", "public_memory"):
        if key not in pub:
            raise ValueError("public input lacks %r" % key)
    if pub["n_steps"] & (pub["n_steps"] - 1):
        raise ValueError("n_steps must be a power of two in proof mode")
    return pub


def write_public_input(path, layout, n_steps, memory_segments, public_memory, rc_min=0, rc_max=0):
    with open(path, "w") as f:
        json.dump({"layout": layout, "n_steps": n_steps, "rc_min": rc_min, "rc_max": rc_max, "memory_segments": memory_segments,
                   "public_memory": [{"address": int(a), "value": hex(int(v)), "page": 0} for a, v in public_memory]}, f)


# ------------------------------------------------------------------ consistency of the three files
def check_consistency(trace, memory, pub):
    """Cheap structural checks a prover makes before anything else: step count, pc / ap / fp inside the program and
    execution segments, public memory cells present with the declared values."""
    if len(trace) != pub["n_steps"]:
        raise ValueError("trace has %d steps, public input says %d" % (len(trace), pub["n_steps"]))
    seg = pub["memory_segments"]
    prog, exe = seg["program"], seg["execution"]
    pc, ap, fp = trace["pc"], trace["ap"], trace["fp"]
    if pc.min() < prog["begin_addr"] or pc.max() >= prog["stop_ptr"]:
        raise ValueError("pc leaves the program segment")
    if ap.min() < exe["begin_addr"] or fp.min() < exe["begin_addr"]:
        raise ValueError("ap / fp below the execution segment")
    if pub["public_memory"]:
        addrs = np.array([c["address"] for c in pub["public_memory"]], dtype=np.uint64)
        want = np.frombuffer(b"".join(int(c["value"], 16).to_bytes(32, "little") for c in pub["public_memory"]),
                             dtype="<u8").reshape(-1, 4)
        if not np.array_equal(memory.get(addrs), want):
            raise ValueError("public memory differs from the memory file")
    return True


# ------------------------------------------------------------------ builtin segments
def pedersen_instances(memory, pub):
    """-> (x, y, result) arrays (n, 4) of the COMPLETE instances of the pedersen builtin segment, and their count"""
    seg = pub["memory_segments"]["pedersen"]
    vals, mask = memory.segment(seg["begin_addr"], seg["stop_ptr"])
    n = vals.shape[0] // PEDERSEN_CELLS
    v = vals[:n * PEDERSEN_CELLS].reshape(n, PEDERSEN_CELLS, 4)
    if not mask[:n * PEDERSEN_CELLS].all():
        raise ValueError("pedersen segment has holes below its stop pointer")
    return v[:, 0].copy(), v[:, 1].copy(), v[:, 2].copy()


def check_pedersen_builtin(memory, pub, ctx=None):
    """Recompute every instance of the Pedersen builtin on the GPU.  Returns the number of instances; raises ValueError
    naming the first instance whose result cell is not pedersen_hash(x, y) (signature.py:296-318)."""
    ctx = ctx or get_context()
    x, y, h = pedersen_instances(memory, pub)
    if x.shape[0] == 0:
        return 0
    got, st = ctx.pedersen_hash2(x, y)
    bad = np.nonzero((st != 0) | (got != h).any(axis=1))[0]
    if bad.size:
        raise ValueError("pedersen builtin instance %d is inconsistent (status %d)" % (int(bad[0]), int(st[bad[0]])))
    return x.shape[0]


def ecdsa_instances(memory, pub):
    seg = pub["memory_segments"]["ecdsa"]
    vals, mask = memory.segment(seg["begin_addr"], seg["stop_ptr"])
    n = vals.shape[0] // ECDSA_CELLS
    if not mask[:n * ECDSA_CELLS].all():
        raise ValueError("ecdsa segment has holes below its stop pointer")
    v = vals[:n * ECDSA_CELLS].reshape(n, ECDSA_CELLS, 4)
    return v[:, 0].copy(), v[:, 1].copy()


def check_ecdsa_builtin(memory, pub, signatures, ctx=None):
    """signatures: {instance index: (r, s)} as the runner's private input carries them.  Every instance of the ECDSA
    builtin must verify (signature.py:217-260 with x-only keys).  Returns the instance count."""
    from ._lib import ints_to_limbs
    ctx = ctx or get_context()
    keys, msgs = ecdsa_instances(memory, pub)
    n = keys.shape[0]
    if n == 0:
        return 0
    if sorted(signatures) != list(range(n)):
        raise ValueError("private input must carry one signature per ecdsa instance")
    r = ints_to_limbs([signatures[i][0] for i in range(n)])
    s = ints_to_limbs([signatures[i][1] for i in range(n)])
    st = ctx.ecdsa_verify(msgs, r, s, keys, None)
    bad = np.nonzero(st != 1)[0]
    if bad.size:
        raise ValueError("ecdsa builtin instance %d does not verify (status %d)" % (int(bad[0]), int(st[bad[0]])))
    return n


def prove_ecdsa_builtin(memory, pub, signatures, n_queries=30, ctx=None):
    """One STARK proof (the ECDSA-builtin AIR, csrc/air_ecdsa.cu) that every instance of the ECDSA builtin segment carries
    a verifying signature: the public input of the proof is exactly the segment's cells -- (message, key x) per instance --
    so a verifier holding the memory file checks `statement["msgs"] / ["keys"]` against it.  signatures: {instance: (r, s)}
    from the runner's private input.  The batch is padded to a power of two (>= 2) by repeating instance 0.
    Returns (proof bytes, log_n, number of real instances)."""
    from ._lib import ints_to_limbs, limbs_to_ints
    from .ecdsa_air import BLOCK, air_inputs
    ctx = ctx or get_context()
    keys_x, msgs = ecdsa_instances(memory, pub)
    n = keys_x.shape[0]
    if n == 0:
        raise ValueError("the ecdsa segment is empty")
    if sorted(signatures) != list(range(n)):
        raise ValueError("private input must carry one signature per ecdsa instance")
    r = ints_to_limbs([signatures[i][0] for i in range(n)])
    s = ints_to_limbs([signatures[i][1] for i in range(n)])
    # the key's y: the builtin cell holds x only; verify() tries both roots (signature.py:229-241) -- keep the one that verifies
    y, st = ctx.get_y_coordinate(keys_x)
    if st.any():
        raise ValueError("ecdsa builtin instance %d: the key is not on the curve" % int(np.nonzero(st)[0][0]))
    ok = ctx.ecdsa_verify(msgs, r, s, keys_x, y) == 1
    yi = limbs_to_ints(y)
    yi = [v if good else FIELD_PRIME - v for v, good in zip(yi, ok)]
    y = ints_to_limbs(yi)
    bad = np.nonzero(ctx.ecdsa_verify(msgs, r, s, keys_x, y) != 1)[0]
    if bad.size:
        raise ValueError("ecdsa builtin instance %d does not verify" % int(bad[0]))
    count = 2
    while count < n:
        count *= 2
    idx = list(range(n)) + [0] * (count - n)
    mi, ri, si, kx = (limbs_to_ints(a) for a in (msgs, r, s, keys_x))
    m_, r_, w_, kx_, ky_ = air_inputs([mi[i] for i in idx], [ri[i] for i in idx], [si[i] for i in idx],
                                      [(kx[i], yi[i]) for i in idx])
    log_n = (count * BLOCK).bit_length() - 1
    trace = ctx.ecdsa_air_trace(log_n, m_, r_, w_, kx_, ky_)
    return ctx.prove_ecdsa(trace, log_n, m_, kx_, n_queries), log_n, n


def load(prefix):
    """(trace, memory, public_input) of `<prefix>_trace.bin`, `<prefix>_memory.bin`, `<prefix>_public_input.json`"""
    trace, memory = read_trace(prefix + "_trace.bin"), read_memory(prefix + "_memory.bin")
    pub = read_public_input(prefix + "_public_input.json")
    check_consistency(trace, memory, pub)
    return trace, memory, pub


assert FIELD_PRIME.bit_length() == 252
