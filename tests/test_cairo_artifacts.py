"""cairo-run artefact ingestion (stark_perpetual_b200/cairo_artifacts.py): the formats are
cairo-lang's (external, cairo_cmake_rules.cmake:72-110 names only the files), so the parser is exercised on files this
test writes itself; the builtin-segment checks run on the GPU against hashes / signatures made by the oracle."""
import random

import numpy as np
import pytest

from oracle import ecdsa as oecdsa
from oracle.params import EC_ORDER, FIELD_PRIME as P
from oracle.pedersen import pedersen_hash
from stark_perpetual_b200 import cairo_artifacts as ca
from stark_perpetual_b200._lib import ints_to_limbs


def _make(tmp_path, n_ped=12, n_sig=3, corrupt=None):
    rng = random.Random(11)
    n_steps = 64
    prog = {"begin_addr": 1, "stop_ptr": 41}
    exe = {"begin_addr": 41, "stop_ptr": 300}
    ped = {"begin_addr": 300, "stop_ptr": 300 + 3 * n_ped}
    ecd = {"begin_addr": 400, "stop_ptr": 400 + 2 * n_sig}
    addr, vals = [], []
    for a in range(prog["begin_addr"], prog["stop_ptr"]):
        addr.append(a); vals.append(rng.randrange(P))
    for a in range(exe["begin_addr"], exe["begin_addr"] + 100):
        addr.append(a); vals.append(rng.randrange(P))
    for i in range(n_ped):
        x, y = rng.randrange(P), rng.randrange(P)
        h = pedersen_hash(x, y)
        if corrupt == ("pedersen", i):
            h ^= 1
        for k, v in enumerate((x, y, h)):
            addr.append(ped["begin_addr"] + 3 * i + k); vals.append(v)
    sigs = {}
    for i in range(n_sig):
        priv, msg = rng.randrange(1, EC_ORDER), rng.randrange(1, 2**250)
        sigs[i] = oecdsa.sign(msg, priv)
        if corrupt == ("ecdsa", i):
            msg ^= 2
        addr.append(ecd["begin_addr"] + 2 * i); vals.append(oecdsa.private_to_stark_key(priv))
        addr.append(ecd["begin_addr"] + 2 * i + 1); vals.append(msg)
    perm = list(range(len(addr)))
    rng.shuffle(perm)                      # the memory file need not be sorted
    prefix = str(tmp_path / "run")
    ca.write_memory(prefix + "_memory.bin", np.array([addr[p] for p in perm], dtype=np.uint64), ints_to_limbs([vals[p] for p in perm]))
    pc = np.array([rng.randrange(prog["begin_addr"], prog["stop_ptr"]) for _ in range(n_steps)], dtype=np.uint64)
    ap = np.arange(n_steps, dtype=np.uint64) + np.uint64(exe["begin_addr"] + 2)
    ca.write_trace(prefix + "_trace.bin", ap, ap - np.uint64(1), pc)
    public_memory = [(a, vals[addr.index(a)]) for a in range(prog["begin_addr"], prog["begin_addr"] + 5)]
    ca.write_public_input(prefix + "_public_input.json", "perpetual_with_bitwise", n_steps,
                          {"program": prog, "execution": exe, "pedersen": ped, "ecdsa": ecd}, public_memory, 0, 65535)
    return prefix, sigs


def test_round_trip_and_consistency(tmp_path):
    prefix, _ = _make(tmp_path)
    trace, memory, pub = ca.load(prefix)
    assert len(trace) == 64 and pub["layout"] == "perpetual_with_bitwise" and len(memory) == 40 + 100 + 36 + 6
    assert int(trace["fp"][3]) == int(trace["ap"][3]) - 1
    x, y, h = ca.pedersen_instances(memory, pub)
    assert x.shape == (12, 4) and (memory.addr[1:] > memory.addr[:-1]).all()
    # structural failures are reported, not ignored
    pub_bad = dict(pub, n_steps=32)
    with pytest.raises(ValueError, match="steps"):
        ca.check_consistency(trace, memory, pub_bad)
    pm = [dict(c) for c in pub["public_memory"]]
    pm[0]["value"] = hex(int(pm[0]["value"], 16) ^ 1)
    with pytest.raises(ValueError, match="public memory"):
        ca.check_consistency(trace, memory, dict(pub, public_memory=pm))
    with open(prefix + "_trace.bin", "ab") as f:
        f.write(b"\0" * 5)
    with pytest.raises(ValueError, match="multiple"):
        ca.read_trace(prefix + "_trace.bin")


@pytest.mark.gpu
def test_builtin_segments_recomputed_on_the_gpu(ctx, tmp_path):
    prefix, sigs = _make(tmp_path)
    _trace, memory, pub = ca.load(prefix)
    assert ca.check_pedersen_builtin(memory, pub, ctx) == 12
    assert ca.check_ecdsa_builtin(memory, pub, sigs, ctx) == 3
    (tmp_path / "b").mkdir(); (tmp_path / "c").mkdir()
    prefix, sigs = _make(tmp_path / "b", corrupt=("pedersen", 7))
    _t, memory, pub = ca.load(prefix)
    with pytest.raises(ValueError, match="instance 7"):
        ca.check_pedersen_builtin(memory, pub, ctx)
    prefix, sigs = _make(tmp_path / "c", corrupt=("ecdsa", 1))
    _t, memory, pub = ca.load(prefix)
    with pytest.raises(ValueError, match="instance 1"):
        ca.check_ecdsa_builtin(memory, pub, sigs, ctx)


@pytest.mark.gpu
def test_ecdsa_builtin_segment_is_proven(ctx, tmp_path):
    """the ECDSA builtin segment of the artefacts -> one proof of the second AIR whose public input is the segment's cells"""
    from oracle import stark
    prefix, sigs = _make(tmp_path, n_sig=3)
    _trace, memory, pub = ca.load(prefix)
    proof, log_n, n = ca.prove_ecdsa_builtin(memory, pub, sigs, ctx=ctx)
    assert (log_n, n) == (10, 3)                       # 3 instances padded to 4 blocks of 256 rows
    st = stark.verify(proof)
    keys, msgs = ca.ecdsa_instances(memory, pub)
    from stark_perpetual_b200._lib import limbs_to_ints
    assert st["air"] == "ecdsa" and st["msgs"][:3] == limbs_to_ints(msgs) and st["keys"][:3] == limbs_to_ints(keys)
    assert st["msgs"][3] == st["msgs"][0] and st["keys"][3] == st["keys"][0]
    (tmp_path / "c").mkdir()
    prefix, sigs = _make(tmp_path / "c", corrupt=("ecdsa", 1))
    _t, memory, pub = ca.load(prefix)
    with pytest.raises(ValueError, match="instance 1"):
        ca.prove_ecdsa_builtin(memory, pub, sigs, ctx=ctx)
