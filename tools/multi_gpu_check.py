#!/usr/bin/env python3
"""torchrun --nproc-per-node G tools/multi_gpu_check.py [log_n]: the sharded prover on G GPUs must emit the
same bytes as the single-GPU spg_prove (computed on rank 0) and the proof must verify under the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import stark_perpetual_b200 as spg  # noqa: E402
from stark_perpetual_b200 import prover  # noqa: E402
from stark_perpetual_b200._lib import limbs_to_ints  # noqa: E402
from conftest import rand_felts  # noqa: E402


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 14
    chain_log = 2
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = spg.Context(local)
    pv = prover.Prover(ctx, rank, world)
    x0 = limbs_to_ints(rand_felts(5, 11))
    ys = rand_felts(5 * ((1 << log_n) >> 9), 12)
    trace = ctx.pedersen_chain_trace(log_n, chain_log, x0, ys)
    n = 1 << log_n
    tr3 = trace.reshape(25, n, 4)
    outs = limbs_to_ints(tr3[[5 * l for l in range(5)], n - 1])
    mine = np.ascontiguousarray(tr3[pv.cyclic_columns()]).reshape(-1, 4)
    proof = pv.prove_cyclic(mine, log_n, chain_log, x0, outs, 30)                 # host buffers: the e2e path
    dev = torch.from_numpy(mine.view(np.int64)).cuda()
    proof_dev = pv.prove_cyclic(None, log_n, chain_log, x0, outs, 30, device_ptr=dev.data_ptr())
    assert proof_dev == proof, "device-pointer and host-pointer sharded proofs differ"
    # the Python-sequenced stage driver (block column deal) must give the same bytes too
    block, outs2 = pv.shard_host_trace(trace, log_n)
    assert outs2 == outs
    assert pv.prove_sharded_device(block, log_n, chain_log, x0, outs, 30) == proof, "Python stage driver differs"
    dist.barrier()
    # the second AIR (ECDSA builtin) through the same sharded sequence: every rank builds the same trace from the same seed
    import random
    from stark_perpetual_b200._lib import ints_to_limbs
    from stark_perpetual_b200.ecdsa_air import air_inputs
    rng = random.Random(5)
    count = n >> 8
    privs = [rng.randrange(1, 1 << 250) for _ in range(count)]
    msgs = [rng.randrange(1, 1 << 251) for _ in range(count)]
    kx, ky, _st = ctx.private_to_stark_key(ints_to_limbs(privs), want_y=True)
    r, s_, _st = ctx.sign(ints_to_limbs(msgs), ints_to_limbs(privs))
    m_, r_, w_, kx_, ky_ = air_inputs(msgs, limbs_to_ints(r), limbs_to_ints(s_), list(zip(limbs_to_ints(kx), limbs_to_ints(ky))))
    etrace = ctx.ecdsa_air_trace(log_n, m_, r_, w_, kx_, ky_)
    emine = np.ascontiguousarray(etrace.reshape(25, n, 4)[pv.cyclic_columns()]).reshape(-1, 4)
    eproof = ctx.prove_ecdsa_sharded(emine, log_n, m_, kx_, 30)
    edev = torch.from_numpy(emine.view(np.int64)).cuda()
    assert ctx.prove_ecdsa_sharded(None, log_n, m_, kx_, 30, device_ptr=edev.data_ptr()) == eproof
    dist.barrier()
    if rank == 0:
        from oracle import stark
        want = ctx.prove(trace, log_n, chain_log, x0, 30)
        stark.verify(proof)
        ewant = ctx.prove_ecdsa(etrace, log_n, m_, kx_, 30)
        assert stark.verify(eproof)["air"] == "ecdsa"
        print("MULTI_GPU_CHECK world=%d log_n=%d bytes=%d %s   ecdsa-air bytes=%d %s" % (
            world, log_n, len(proof), "MATCH" if proof == want else "MISMATCH", len(eproof), "MATCH" if eproof == ewant else "MISMATCH"))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
