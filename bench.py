#!/usr/bin/env python3
"""bench.py -- the driver-facing benchmark of the proving hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun by the driver)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one full proof (trace LDE blowup 8 -> Merkle -> AIR composition -> chunk LDE -> Merkle -> OODS ->
DEEP -> FRI -> queries) of one synthetic but VALID 2^20-row, 25-column Pedersen hash-chain trace -- the
configuration BASELINE.json's metric is quoted on ("2^20-step trace", configs[2] plus the FRI/Merkle stages of
configs[3]).  The trace is FIXED as N grows (strong scaling).

Prints ONE JSON line on rank 0.  `value` = algorithmic 252-bit field multiplications per second over the
whole job with the trace resident in HBM; `proof_gen_s` = seconds per proof; `e2e` = the same through the
host-buffer C-ABI call spg_prove (H2D of the 840 MB trace and D2H of the proof inside the timed region).
`--impl reference` times the CPU restatement (oracle/c, OpenMP on all host cores) on a bounded sample -- the
reference repository itself contains no prover to time, so kind = "port".
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "field_mul_per_s_2^20_trace_proof"
UNIT = "252-bit field-mul/s"
STAGES = ["lde_trace", "merkle_trace", "air_composition", "lde_chunks", "merkle_chunks", "oods_eval", "deep_quotient",
          "fri", "queries"]


def workload_cfg(args):
    return {"log_n": args.log_n, "n_cols": 25, "log_blowup": 3, "chain_log": args.chain_log, "n_queries": args.queries}


def lde_muls(n_cols, log_n, blowup=8):
    """butterflies (N/2 log2 N per transform) + one multiplication per point for the
    scaling / coset shift, per column."""
    n = 1 << log_n
    return n_cols * n * ((1 + blowup) * log_n / 2.0 + (1 + blowup))


def proof_muls(cfg):
    """Algorithmic multiplication count of one proof, stage by stage (DESIGN.md "Work per proof")."""
    n, ln = 1 << cfg["log_n"], cfg["log_n"]
    st = {
        "lde_trace": lde_muls(25, ln),
        "air_composition": 4 * n * (5 * 21 + 8) + n * 12,     # 21 per lane + 8 zerofier products; chunk split
        "lde_chunks": lde_muls(4, ln),
        "oods_eval": 54 * n,
        "deep_quotient": 8 * n * (54 + 3) + 3 * 8 * n * 5,    # quotient + batched inversion (5 per point)
        "fri": 15 * n * 8 / 7.0,                              # fold-by-8 layers, geometric
    }
    return sum(st.values()), st


def host_trace_random(n_cols, log_n, seed):
    from conftest import rand_felts
    return rand_felts(n_cols << log_n, seed)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx, self.marks = [], None, gpu_index, []

    def mark(self):
        self.marks.append(len(self.rows))

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = self.rows
        if len(self.marks) >= 2 and self.marks[1] > self.marks[0]:
            rows = self.rows[self.marks[0]:self.marks[1] + 1]      # the timed region only
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 9 for i in range(4) if r[5 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_sample(cfg):
    """Time the C oracle (OpenMP) on a bounded sample of the proof's dominant stage: the blowup-8 LDE of as
    many 2^log_n columns as there are host threads (one per thread).  Returns (mul/s, cores, text, seconds)."""
    from oracle import clib
    cores = clib.use_all_cores()
    sample_cols = max(1, min(25, cores))
    tr = host_trace_random(sample_cols, cfg["log_n"], 4242)
    t0 = time.perf_counter()
    clib.lde(tr, cfg["log_n"], sample_cols, cfg["log_blowup"])
    dt = time.perf_counter() - t0
    muls = lde_muls(sample_cols, cfg["log_n"])
    return muls / dt, cores, ("LDE stage (%.0f%% of a proof's multiplications) of %d of the 25 columns of the 2^%d trace, "
                              "blowup 8, plain-C oracle with OpenMP on %d threads, %.1f s") % (
        100 * lde_muls(25, cfg["log_n"]) / proof_muls(cfg)[0], sample_cols, cfg["log_n"], cores, dt), dt


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = workload_cfg(args)
    vals, dts, desc, cores = [], [], "", 1
    for i in range(args.warmup + args.steps):
        v, cores, desc, dt = cpu_sample(cfg)
        if i >= args.warmup:
            vals.append(v); dts.append(dt)
        if sum(dts) > 150:      # keep the whole run within a few minutes
            break
    v = float(np.mean(vals))
    total, _ = proof_muls(cfg)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(dts)),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u256 (4x u64 Montgomery)",
            "data": "synthetic", "config": dict(workload="stark_proof_2^%d_x25_blowup8" % cfg["log_n"], **cfg),
            "proof_gen_s_extrapolated": total / v,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference repository has no prover (SURVEY.md section 0); this is the repo's own plain-C "
                    "restatement, all host threads, on a bounded sample (one step = the sample)"}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------- orders workload
def orders_arm(args):
    """BASELINE.json configs[4]: batch-verify synthetic perpetual limit orders (packing + 4 Pedersen hashes + STARK-curve
    ECDSA per order), sharded across the ranks as independent units: no collective on the data path (weak in the
    sense of SURVEY.md section 8e: fixed total of --orders, contiguous n/G split)."""
    import torch
    import torch.distributed as dist
    import stark_perpetual_b200 as spg
    from conftest import rand_felts
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = spg.Context(local)
    n_total = args.orders
    n = n_total // world
    g = np.random.Generator(np.random.PCG64(1005 + rank))
    orders = {"asset_id_synthetic": rand_felts(n, 41 + rank), "asset_id_collateral": rand_felts(n, 142 + rank),
              "asset_id_fee": rand_felts(n, 243 + rank), "is_buying_synthetic": g.integers(0, 2, n, dtype=np.uint8)}
    orders["asset_id_synthetic"][:, 2:] = 0
    for f in ("asset_id_collateral", "asset_id_fee"):
        orders[f][:, 3] &= np.uint64((1 << 58) - 1)
    for f in ("amount_synthetic", "amount_collateral", "max_amount_fee", "position_id"):
        orders[f] = g.integers(0, 2**64, n, dtype=np.uint64)
    for f in ("nonce", "expiration_timestamp"):
        orders[f] = g.integers(0, 2**32, n, dtype=np.uint32)
    r, s_ = rand_felts(n, 32 + rank), rand_felts(n, 33 + rank)
    for a in (r, s_):
        a[:, 3] &= np.uint64(0x07ffffffffffffff)
    keys, _st = ctx.private_to_stark_key(rand_felts(64, 7))
    px = np.tile(keys, (n // 64 + 1, 1))[:n]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        ctx.limit_order_verify(orders, r, s_, px)
    launches_before = ctx.launch_count
    barrier()
    t0 = time.perf_counter()
    kernel_ms = 0.0
    for _ in range(args.steps):
        st = ctx.limit_order_verify(orders, r, s_, px)       # host buffers in, statuses out: the e2e call
        kernel_ms += ctx.last_kernel_ms
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    t = torch.tensor([kernel_ms / args.steps, wall_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    k_ms, w_ms = float(t[0].item()), float(t[1].item())
    if rank == 0:
        per_order_in = 3 * 32 + 1 + 4 * 8 + 2 * 4 + 3 * 32
        line = {"metric": "limit_orders_verified_per_s", "value": n * world / (k_ms * 1e-3), "unit": "orders/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": k_ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u256 (8x u32 limbs, Montgomery)", "data": "synthetic",
                "config": {"workload": "limit_order_verify_batch", "orders": n * world, "per_gpu": n,
                           "parallelism": "independent units, contiguous n/G split, no collective"},
                "gpu_launches": (ctx.launch_count - launches_before),
                "e2e": {"value": n * world / (w_ms * 1e-3), "unit": "orders/s", "ms_per_step": w_ms,
                        "h2d_bytes_per_step": n * world * per_order_in, "d2h_bytes_per_step": n * world},
                "status_counts": {int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def multi_gpu_aux(ctx, pv, args, rank, world, dev, barrier):
    """BASELINE.json configs[3] and [4] at N > 1, measured after the timed region: a full 2^22-row proof sharded over the
    ranks (verified by the oracle on rank 0) and the 65536-order batch verification split n/N per rank (independent
    units, no collective).  Every rank takes part; returns the object on rank 0."""
    import torch
    import torch.distributed as dist
    from conftest import rand_felts
    from stark_perpetual_b200._lib import limbs_to_ints
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import aux_bench
    out = {}

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    # ---- configs[3]: full proof of a 2^22-row trace
    log_n = args.aux_log_n
    n = 1 << log_n
    x0 = limbs_to_ints(rand_felts(5, 2203))
    trace = ctx.pedersen_chain_trace(log_n, args.chain_log, x0, rand_felts(5 * (n >> 9), 2204))
    tr3 = trace.reshape(25, n, 4)
    outs = limbs_to_ints(tr3[[5 * l for l in range(5)], n - 1])
    mine = torch.from_numpy(np.ascontiguousarray(tr3[pv.cyclic_columns()]).view(np.int64)).to(dev)
    del trace, tr3
    proof = None
    for _ in range(2):
        proof = pv.prove_cyclic(None, log_n, args.chain_log, x0, outs, args.queries, device_ptr=mine.data_ptr())
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.current_stream()
    e0.record(stream)
    for _ in range(3):
        proof = pv.prove_cyclic(None, log_n, args.chain_log, x0, outs, args.queries, device_ptr=mine.data_ptr())
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / 3)
    cfg = dict(workload_cfg(args), log_n=log_n)
    row = {"log_n": log_n, "n_gpus": world, "ms": ms, "field_mul_per_s": proof_muls(cfg)[0] / (ms * 1e-3),
           "proof_bytes": len(proof), "proof_sha256": hashlib.sha256(proof).hexdigest()}
    if rank == 0 and not args.no_verify:
        from oracle import stark as ostark
        ostark.verify(proof)
        row["verified_by_oracle"] = True
    out["cfg3_proof_2^%d" % log_n] = row
    del mine
    torch.cuda.empty_cache()
    # ---- configs[4]: the order batch split across the ranks
    n_total = args.orders
    n_loc = n_total // world
    orders, r, s_, px, expect, _bad, _msgs, _nk, _sb = aux_bench.signed_orders(ctx, n_loc, 1005 + rank)
    ctx.limit_order_verify(orders, r, s_, px)
    barrier()
    t0 = time.perf_counter()
    st = ctx.limit_order_verify(orders, r, s_, px)
    k_ms = ctx.last_kernel_ms
    barrier()
    wall = max_over_ranks(1e3 * (time.perf_counter() - t0))
    k_ms = max_over_ranks(k_ms)
    ok = max_over_ranks(0.0 if np.array_equal(st, expect) else 1.0) == 0.0
    out["cfg4_orders_split"] = {"n": n_loc * world, "per_gpu": n_loc, "n_gpus": world, "ms": k_ms, "e2e_ms": wall,
                                "orders_per_s": n_loc * world / (k_ms * 1e-3), "e2e_orders_per_s": n_loc * world / (wall * 1e-3),
                                "statuses_as_expected": ok, "parallelism": "independent units, contiguous n/G split, no collective"}
    return out if rank == 0 else None


# ----------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--log-n", dest="log_n", type=int, default=20)
    ap.add_argument("--chain-log", dest="chain_log", type=int, default=2)
    ap.add_argument("--queries", type=int, default=30)
    ap.add_argument("--workload", default="proof", choices=["proof", "orders"],
                    help="proof: the headline 2^20-row proof (default); orders: BASELINE configs[4] batch verification")
    ap.add_argument("--orders", type=int, default=65536)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the aux object (BASELINE configs 0, 1, 4 after the timed region)")
    ap.add_argument("--aux-log-n", dest="aux_log_n", type=int, default=22, help="N > 1: size of the configs[3] proof in aux")
    ap.add_argument("--no-verify", action="store_true", help="skip the oracle verification of one proof")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    if args.workload == "orders":
        return orders_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import stark_perpetual_b200 as spg
    from stark_perpetual_b200 import prover

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = workload_cfg(args)
    log_n, n = cfg["log_n"], 1 << cfg["log_n"]
    ctx = spg.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    pv = prover.Prover(ctx, rank=rank, world=world)

    # ---- synthetic but valid trace (seed 1003, SURVEY.md section 8(d) cfg 3a): witness generated on the device
    rng = np.random.Generator(np.random.PCG64(1003))
    from conftest import rand_felts
    inst = n >> 9
    x0_limbs = rand_felts(5, 1003)
    ys_limbs = rand_felts(5 * inst, 1004)
    from stark_perpetual_b200._lib import limbs_to_ints
    x0 = limbs_to_ints(x0_limbs)
    trace_host = ctx.pedersen_chain_trace(log_n, cfg["chain_log"], x0, ys_limbs)          # (25 N, 4) uint64
    # every rank pins and uploads only its own columns (world = 1: the whole trace; world > 1: the cyclic deal
    # rank, rank + world, ... of csrc/sharded.cu)
    tr3 = trace_host.reshape(25, n, 4)
    outs = limbs_to_ints(tr3[[5 * l for l in range(5)], n - 1])
    my = tr3 if world == 1 else tr3[pv.cyclic_columns()]
    if world > 1:
        pv.setup_comm()
    pinned = torch.from_numpy(np.ascontiguousarray(my).view(np.int64)).pin_memory()
    trace = torch.empty_like(pinned, device=dev)
    trace.copy_(pinned)
    torch.cuda.synchronize()
    launches_before = ctx.launch_count

    def step():
        if world == 1:
            return pv.prove_device(trace.data_ptr(), log_n, cfg["chain_log"], x0, cfg["n_queries"])
        return pv.prove_cyclic(None, log_n, cfg["chain_log"], x0, outs, cfg["n_queries"], device_ptr=trace.data_ptr())

    def step_e2e():
        if world == 1:
            return pv.prove_host(pinned.numpy().view(np.uint64), log_n, cfg["chain_log"], x0, cfg["n_queries"])
        return pv.prove_cyclic(pinned.numpy().view(np.uint64), log_n, cfg["chain_log"], x0, outs, cfg["n_queries"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    proof = None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # nvidia-smi takes a moment to deliver its first row: started before the warm-up,
    for _ in range(args.warmup):  # only the rows of the timed region are kept (mark / stop below)
        proof = step()
    barrier()
    launches_per_step = (ctx.launch_count - launches_before) // args.warmup
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = np.zeros(len(STAGES))
    barrier()
    sampler.mark()
    e0.record(stream)
    for _ in range(args.steps):
        proof = step()
        stage_ms += np.array([ctx.stage_ms(i) for i in range(len(STAGES))])
    e1.record(stream)
    barrier()
    sampler.mark()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    stage_ms /= args.steps

    # ---- e2e: host trace through the C-ABI (pinned host memory): H2D + proof + D2H of the proof bytes
    e2e = None
    if not args.no_e2e:
        step_e2e()
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            pr2 = step_e2e()
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"ms_per_step": float(t.item()) / args.steps, "h2d_bytes_per_step": 25 * n * 32,   # summed over ranks
               "d2h_bytes_per_step": len(pr2) if pr2 is not None else 0}

    line = None
    if rank == 0:
        muls, per_stage = proof_muls(cfg)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        line = {"metric": METRIC, "value": muls / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u256 (8x u32 limbs, Montgomery)",
                "data": "synthetic",
                "config": dict(workload="stark_proof_2^%d_x25_blowup8" % log_n, **cfg,
                               l2="working set %.1f GB per step, far above the 126 MB L2; no flush needed" % (
                                   (1 + 8) * 29 * n * 32 / 1e9),
                               parallelism=pv.parallelism()),
                "proof_gen_s": ms_per_step * 1e-3, "proof_bytes": len(proof) if proof else None,
                "proof_sha256": hashlib.sha256(proof).hexdigest() if proof else None,
                "stage_ms": {s: round(float(v), 4) for s, v in zip(STAGES, stage_ms)},
                "algorithmic_muls_per_step": muls,
                "gpu_launches": launches_per_step * args.steps, "clocks": clocks}
        # dominant kernel: k_ntt_tile (all launches of the two LDE stages).  SURVEY.md section 8(d): the compulsory HBM traffic
        # of a size-N transform is 64 N bytes (read once, write once), of the blowup-8 LDE 32 N C (1 + 8); a transform of
        # 2^log_n > 2^10 points takes two launches (passes), so one launch is charged half a transform: 32 N bytes per
        # column.  `frac` is that figure; `per_pass` charges every launch the 64 N bytes it really moves (two-pass
        # algorithm: each pass reads and writes its columns once), which is what ncu's DRAM counters see (`traffic`).
        ntt_ms = float(stage_ms[0] + stage_ms[3])
        if ntt_ms > 0 and world == 1:
            passes = 2 if log_n > 10 else 1
            n_launch = (1 + 8) * 2 * passes
            alg_bytes = (1 + 8) * 29 * n * 64.0                    # = 32 N C (1 + B) + the 4 chunk columns, read + write
            pass_bytes = alg_bytes * passes
            ach = alg_bytes / (ntt_ms * 1e-3) / 1e9
            traffic, traffic_src = None, None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["k_ntt_tile"]
                if log_n == 20:
                    traffic = tj["dram_bytes_per_launch"]
                    traffic_src = "static: profiles/ncu_traffic.json (%s)" % tj.get("source", "ncu --set full capture of this kernel")
            except (OSError, KeyError, ValueError):
                pass
            ntt_muls = per_stage["lde_trace"] + per_stage["lde_chunks"]
            # the binding resource: IMAD.WIDE issues at one warp instruction per 4 cycles per SM sub-partition
            # (32 lanes/clk/SM, measured: profiles/r1g_microbench_imad.json); a multiplication needs 64 of them
            sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
            imad_peak = 148 * 32 * sm_mhz * 1e6
            line["roofline"] = {"bound": "hbm", "kernel": "k_ntt_tile", "achieved": ach, "peak": peak, "unit": "GB/s",
                                "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src,
                                "algorithmic_bytes_per_launch": alg_bytes / n_launch,
                                "per_pass": {"bytes_per_launch": pass_bytes / n_launch,
                                             "achieved": pass_bytes / (ntt_ms * 1e-3) / 1e9,
                                             "frac": pass_bytes / (ntt_ms * 1e-3) / 1e9 / peak},
                                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650",
                                "launches_per_step": n_launch, "avg_launch_ms": ntt_ms / n_launch,
                                "share_of_step": ntt_ms / ms_per_step,
                                "field_mul_per_s": ntt_muls / (ntt_ms * 1e-3),
                                "executed_muls_per_element": {"inverse": 10.5, "per_coset": 10.5,
                                                              "note": "8.5 butterfly + 1 diagonal + 1 scale / coset-shift "
                                                                      "multiplication per element per transform (DESIGN.md section 3)"},
                                "int_pipe": {"binding": "fmaheavy (IMAD.WIDE)", "imad_wide_per_s_peak": imad_peak,
                                             "field_mul_per_s_ceiling": imad_peak / 64.0,
                                             "frac_of_ceiling": ntt_muls / (ntt_ms * 1e-3) / (imad_peak / 64.0),
                                             "executed_frac_of_ceiling": (1 + 8) * 29 * n * 10.5 / (ntt_ms * 1e-3) / (imad_peak / 64.0)},
                                "note": "HBM fraction reported as the contract asks (SURVEY 8(d) compulsory bytes); the kernel is "
                                        "bound by the integer multiply pipe, not by HBM (DESIGN.md section 2: 64 IMAD.WIDE per "
                                        "252-bit multiplication at 32 lanes/clk/SM)"}
        if e2e:
            line["e2e"] = {"value": muls / (e2e["ms_per_step"] * 1e-3), "unit": UNIT,
                           "h2d_bytes_per_step": e2e["h2d_bytes_per_step"], "d2h_bytes_per_step": e2e["d2h_bytes_per_step"],
                           "ms_per_step": e2e["ms_per_step"], "proof_gen_s": e2e["ms_per_step"] * 1e-3}
        if not args.no_verify and proof is not None:
            from oracle import stark as ostark
            t0 = time.perf_counter()
            ostark.verify(proof)
            line["verified_by_oracle"] = True
            line["verify_s"] = round(time.perf_counter() - t0, 3)
        if world == 1 and not args.no_cpu:
            v, cores, desc, _dt = cpu_sample(cfg)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                                    "proof_gen_s_extrapolated": muls / v}
        if world == 1 and not args.no_aux:
            # BASELINE.json configs[0], [1], [4] on the same record, measured after the timed region (tools/aux_bench.py)
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import aux_bench
            try:
                line["aux"] = aux_bench.measure(ctx, (clocks or {}).get("sm_mhz") or 1965.0, with_reference=not args.no_cpu,
                                                n_orders=args.orders, hbm_gbs=peak)
            except Exception as e:          # the headline line must survive a failure of the side measurements
                line["aux"] = {"error": repr(e)[:500]}
            try:                            # the second AIR (ECDSA builtin) at the headline trace size: 4096 signatures, one proof
                line["aux"]["ecdsa_air_2^%d" % args.log_n] = aux_bench.ecdsa_air(ctx, args.log_n, args.queries)
            except Exception as e:
                line["aux"]["ecdsa_air_error"] = repr(e)[:500]
    if world > 1 and not args.no_aux:
        # BASELINE.json configs[3] (2^22 proof) and [4] (order batch split) at this N, after the timed region.  A watchdog
        # makes sure the headline line is printed even if a rank gets stuck in a collective of the side measurement.
        def give_up():
            if rank == 0:
                line["aux"] = {"error": "multi-GPU aux timed out"}
                print(json.dumps(line), flush=True)
            os._exit(0)
        dog = threading.Timer(300.0, give_up)
        dog.daemon = True
        dog.start()
        try:
            aux_multi = multi_gpu_aux(ctx, pv, args, rank, world, dev, barrier)
        except Exception as e:
            aux_multi = {"error": repr(e)[:500]}
        dog.cancel()
        if rank == 0:
            line["aux"] = aux_multi
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
