"""One rank of the CPU (gloo) multi-process run of the sharded prover driver; see test_sharded_gloo.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cpu_stage_backend import CpuBackend  # noqa: E402
from oracle import stark  # noqa: E402
from stark_perpetual_b200 import prover  # noqa: E402
from stark_perpetual_b200._lib import ints_to_limbs  # noqa: E402


def main():
    out_path, log_n, chain_log, nq = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    import random
    rng = random.Random(123)
    inst = (1 << log_n) // 512
    x0 = [rng.randrange(stark.P) for _ in range(5)]
    ys = [[rng.randrange(stark.P) for _ in range(inst)] for _ in range(5)]
    cols, outs = stark.gen_trace(log_n, chain_log, x0, ys)
    per = -(-25 // world)
    c0, c1 = min(25, rank * per), min(25, (rank + 1) * per)
    be = CpuBackend()
    n = 1 << log_n
    block = be.upload(ints_to_limbs([v for c in cols[c0:c1] for v in c]).reshape(c1 - c0, n, 4)) if c1 > c0 else be.felts(1, n)
    proof = prover.prove_sharded(be, prover.TorchComm(rank, world), block, log_n, chain_log, x0, outs, nq)
    if rank == 0:
        want = stark.prove_trace(log_n, chain_log, x0, outs, cols, n_queries=nq)
        stark.verify(proof, min_queries=nq)
        with open(out_path, "w") as f:
            f.write("OK" if proof == want else "MISMATCH %d %d" % (len(proof), len(want)))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
