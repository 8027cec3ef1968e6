"""Host-side driver of the STARK prover (the role Stone's `cpu_air_prover` CLI plays next to cairo-run's
artefacts, src/starkware/cairo/lang/cairo_cmake_rules.cmake:72-110 -- the reference itself has no prover).

    pv = Prover(ctx)                       # one GPU
    trace = pv.witness(log_n, chain_log, x0, ys)
    proof = pv.prove_host(trace, log_n, chain_log, x0)
    # oracle/stark.py verify(proof) is the independent checker used by the tests

Multi-GPU (one process per GPU, torch.distributed for the rendezvous): see `Prover(ctx, rank, world)` and
DESIGN.md "Multi-GPU".
"""
import numpy as np

from ._lib import ints_to_limbs


class Prover:
    def __init__(self, ctx, rank=0, world=1):
        self.ctx, self.rank, self.world = ctx, rank, world
        if world > 1:
            self._init_comm()

    # ---- single GPU ------------------------------------------------------------------------------
    def witness(self, log_n, chain_log, x0, ys):
        """ys: list of 5 lists of ints (or a (5 * inst, 4) uint64 array).  Returns the (25 N, 4) trace."""
        if not isinstance(ys, np.ndarray):
            ys = ints_to_limbs([v for lane in ys for v in lane])
        return self.ctx.pedersen_chain_trace(log_n, chain_log, x0, ys)

    def prove_host(self, trace, log_n, chain_log, x0, n_queries=30):
        if self.world > 1:
            return self._prove_sharded(trace, None, log_n, chain_log, x0, n_queries)
        return self.ctx.prove(trace, log_n, chain_log, x0, n_queries)

    def prove_device(self, trace_ptr, log_n, chain_log, x0, n_queries=30):
        if self.world > 1:
            return self._prove_sharded(None, trace_ptr, log_n, chain_log, x0, n_queries)
        return self.ctx.prove(None, log_n, chain_log, x0, n_queries, device_ptr=trace_ptr)

    def parallelism(self):
        if self.world == 1:
            return "1 GPU"
        return "%d GPUs: rank 0 proves, the others idle (sharded prover not built yet)" % self.world

    # ---- multi GPU -------------------------------------------------------------------------------
    def _init_comm(self):
        pass

    def _prove_sharded(self, trace, trace_ptr, log_n, chain_log, x0, n_queries):
        if self.rank != 0:
            return None
        if trace_ptr is not None:
            return self.ctx.prove(None, log_n, chain_log, x0, n_queries, device_ptr=trace_ptr)
        return self.ctx.prove(trace, log_n, chain_log, x0, n_queries)
