"""GPU parity at the HEADLINE size (BASELINE.json configs[2]/[3] shape: 2^20 rows x 25 columns, blowup 8, 30 queries,
the exact trace bench.py proves): every stage's values at 4096 sampled points against the oracle (tests/stage_probe.py,
SURVEY.md section 8(d) cfg-3a), the proof accepted by the independent verifier, byte-identical between the stage
driver and the monolithic spg_prove, and -- with two or more GPUs visible -- byte-identical across world sizes."""
import hashlib
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import rand_felts
from oracle import stark
from stark_perpetual_b200._lib import limbs_to_ints

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def bench_inputs(log_n):
    """the inputs bench.py uses (seeds 1003 / 1004)"""
    x0 = limbs_to_ints(rand_felts(5, 1003))
    ys = rand_felts(5 * ((1 << log_n) >> 9), 1004)
    return x0, ys


@pytest.mark.parametrize("log_n,chain_log,n_groups", [(12, 1, 64), (20, 2, 512)])
def test_stage_values_at_sampled_points(ctx, log_n, chain_log, n_groups):
    from stage_probe import StageChecker
    from stark_perpetual_b200 import prover
    x0, ys = bench_inputs(log_n)
    trace = ctx.pedersen_chain_trace(log_n, chain_log, x0, ys)
    pv = prover.Prover(ctx)
    block, outs = pv.shard_host_trace(trace, log_n)
    chk = StageChecker(pv.be, trace, log_n, chain_log, x0, outs, n_groups=n_groups)
    proof = pv.prove_sharded_device(block, log_n, chain_log, x0, outs, n_queries=30, probe=chk)
    chk.assert_complete()
    assert len(chk.points) == 8 * n_groups
    # the same bytes as the monolithic C++ prover (what bench.py times), accepted by the independent verifier
    want = ctx.prove(trace, log_n, chain_log, x0, n_queries=30)
    assert hashlib.sha256(proof).hexdigest() == hashlib.sha256(want).hexdigest()
    st = stark.verify(want)
    assert st["x0"] == x0 and st["outs"] == outs and st["n_queries"] == 30
    print("proof_sha256[2^%d] = %s" % (log_n, hashlib.sha256(want).hexdigest()))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_gpu_proof_is_byte_identical():
    """torchrun --nproc-per-node 2 of tools/multi_gpu_check.py: the sharded proof equals the 1-GPU proof byte for byte."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "multi_gpu_check.py"), "16"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "MATCH" in out.stdout and "MISMATCH" not in out.stdout, out.stdout


@pytest.mark.parametrize("log_n,chain_log", [(10, 1), (16, 2)])
def test_cpp_sharded_prover_world1_equals_monolithic(ctx, log_n, chain_log):
    """csrc/sharded.cu with a one-rank communicator (no NCCL involved): the cyclic column pipeline, chunked coset evaluation,
    root / query assembly must reproduce spg_prove byte for byte, from host and from device buffers."""
    import torch
    x0, ys = bench_inputs(log_n)
    trace = ctx.pedersen_chain_trace(log_n, chain_log, x0, ys)
    n = 1 << log_n
    outs = limbs_to_ints(trace.reshape(25, n, 4)[[5 * l for l in range(5)], n - 1])
    want = ctx.prove(trace, log_n, chain_log, x0, n_queries=30)
    ctx.comm_init(0, 1)
    assert ctx.prove_sharded(trace, log_n, chain_log, x0, outs, 30) == want
    dev = torch.from_numpy(trace.view(np.int64)).cuda()
    assert ctx.prove_sharded(None, log_n, chain_log, x0, outs, 30, device_ptr=dev.data_ptr()) == want


def test_maximum_trace_size_2_23():
    """The largest trace the interface admits (log_n = 23: 8.4 M rows x 25 columns = 6.7 GB, 54 GB extended; three-pass
    transforms): witness on the device, one proof, accepted by the oracle verifier.  Runs in its own process (its ~100 GB of
    device memory are released afterwards); skipped where the device or the host cannot hold it."""
    import torch
    free, _total = torch.cuda.mem_get_info()
    if free < 130e9:
        pytest.skip("needs ~100 GB of free device memory")
    code = r'''
import sys
sys.path.insert(0, "tests")
from test_gpu_headline import bench_inputs
from oracle import stark
import stark_perpetual_b200 as spg
ctx = spg.Context(0)
x0, ys = bench_inputs(23)
trace = ctx.pedersen_chain_trace(23, 2, x0, ys)
proof = ctx.prove(trace, 23, 2, x0)
print("PROOF_MS %.1f" % ctx.last_kernel_ms)
st = stark.verify(proof)
assert st["log_n"] == 23 and st["x0"] == x0
print("VERIFIED", len(proof))
'''
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, cwd=ROOT)
    if out.returncode != 0 and ("out of memory" in out.stderr or "MemoryError" in out.stderr):
        pytest.skip("not enough memory for the 2^23 trace: " + out.stderr[-200:])
    assert out.returncode == 0 and "VERIFIED" in out.stdout, out.stdout[-500:] + out.stderr[-1500:]
    print(out.stdout.strip())
