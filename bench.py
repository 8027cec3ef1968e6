#!/usr/bin/env python3
"""bench.py -- the driver-facing benchmark of the proving hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun by the driver)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one synthetic 2^20-row, 25-column trace (BASELINE.json
configs[2], the configuration the metric is quoted on).  The trace is FIXED as N grows (strong scaling):
columns are sharded for interpolation, one NCCL all-gather exchanges the coefficient columns, cosets are
sharded for evaluation.

Prints ONE JSON line on rank 0.  `value` = algorithmic 252-bit field multiplications per second over the
whole job with inputs resident in HBM; `e2e` = the same through the host-buffer C-ABI call (H2D of the
trace and D2H of the result inside the timed region).  `--impl reference` times the CPU restatement
(oracle/c, OpenMP on all host cores) -- the reference repository itself contains no prover to time
(SURVEY.md section 0), so kind = "port".
"""
import argparse
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

METRIC = "field_mul_per_s_2^20_trace_prove_path"
UNIT = "252-bit field-mul/s"


def workload_cfg(args):
    return {"log_n": args.log_n, "n_cols": args.cols, "log_blowup": args.log_blowup}


def algorithmic_muls(cfg, stages):
    """butterflies (N/2 log2 N per transform) + one multiplication per point for
    scaling / coset shift, per column; plus the counted multiplications of the later stages."""
    n, c, b, ln = 1 << cfg["log_n"], cfg["n_cols"], 1 << cfg["log_blowup"], cfg["log_n"]
    total = c * n * ((1 + b) * ln / 2.0 + (1 + b))
    for s in stages:
        total += s
    return total


def host_trace(cfg, seed):
    from conftest import rand_felts
    return rand_felts(cfg["n_cols"] << cfg["log_n"], seed)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_sample(cfg, cores_cols=None):
    """Time the C oracle (OpenMP) on a bounded sample: as many columns of the same 2^log_n trace as there
    are host threads (one column per thread), full blowup.  Returns (muls_per_s, cores, description)."""
    from oracle import clib
    cores = clib.num_threads()
    sample_cols = min(cfg["n_cols"], cores_cols or cores)
    sub = dict(cfg, n_cols=sample_cols)
    tr = host_trace(sub, 4242)
    t0 = time.perf_counter()
    clib.lde(tr, cfg["log_n"], sample_cols, cfg["log_blowup"])
    dt = time.perf_counter() - t0
    muls = algorithmic_muls(sub, [])
    return muls / dt, cores, "%d of %d columns of the 2^%d trace, blowup %d, C oracle (OpenMP, %d threads), %.1f s" % (
        sample_cols, cfg["n_cols"], cfg["log_n"], 1 << cfg["log_blowup"], cores, dt), dt


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = workload_cfg(args)
    vals, dts = [], []
    for i in range(args.warmup + args.steps):
        v, cores, desc, dt = cpu_sample(cfg)
        if i >= args.warmup:
            vals.append(v); dts.append(dt)
        if sum(dts) > 150:      # keep the whole run within a few minutes
            break
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(dts)),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u256 (4x u64 Montgomery)",
            "data": "synthetic", "config": dict(workload="lde_2^%d_x%d_blowup%d" % (cfg["log_n"], cfg["n_cols"], 1 << cfg["log_blowup"]),
                                                  **cfg),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference repository has no prover (SURVEY.md section 0); this is the repo's own plain-C "
                    "restatement of the same stages, all host threads"}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--log-n", dest="log_n", type=int, default=20)
    ap.add_argument("--cols", type=int, default=25)
    ap.add_argument("--log-blowup", dest="log_blowup", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import stark_perpetual_b200 as spg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = workload_cfg(args)
    n, C, B = 1 << cfg["log_n"], cfg["n_cols"], 1 << cfg["log_blowup"]
    assert B % world == 0, "blowup must be a multiple of the GPU count"
    ctx = spg.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    # column shard of this rank for phase A (contiguous blocks; the last ranks may hold one column less)
    per = (C + world - 1) // world
    c0, c1 = min(C, rank * per), min(C, (rank + 1) * per)
    my_cols = c1 - c0
    cosets = B // world
    full = host_trace(cfg, 1003)                      # seed of SURVEY.md section 8(d) cfg 3a
    host = torch.from_numpy(full.view(np.int64)).reshape(C, n, 4)
    pinned = host[c0:c1].contiguous().pin_memory() if my_cols else None
    trace = torch.empty((max(my_cols, 1), n, 4), dtype=torch.int64, device=dev)
    if my_cols:
        trace[:my_cols].copy_(pinned)
    coeffs = torch.empty((per * world, n, 4), dtype=torch.int64, device=dev)   # padded to equal shards
    out = torch.empty((cosets, C, n, 4), dtype=torch.int64, device=dev)
    launches_before = ctx.launch_count

    def step():
        if world == 1:
            ctx.lde_device(trace.data_ptr(), cfg["log_n"], C, cfg["log_blowup"], out.data_ptr(), sync=False)
        else:
            mine = coeffs[rank * per:(rank + 1) * per]
            if my_cols:
                ctx.lde_coeffs_device(trace.data_ptr(), cfg["log_n"], my_cols, mine.data_ptr(), sync=False)
            dist.all_gather_into_tensor(coeffs, mine)           # the single exchange step
            ctx.lde_cosets_device(coeffs.data_ptr(), cfg["log_n"], C, cfg["log_blowup"], rank * cosets, cosets,
                                  out.data_ptr(), sync=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches_per_step = (ctx.launch_count - launches_before) // args.warmup
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    # dominant kernel (k_ntt_pass) live timing: one synchronous step, stage events inside the library
    if world == 1:
        ctx.lde_device(trace.data_ptr(), cfg["log_n"], C, cfg["log_blowup"], out.data_ptr(), sync=True)
        ntt_ms = ctx.stage_ms(0) + ctx.stage_ms(1)
        n_pass_launches = launches_per_step
    else:
        ntt_ms, n_pass_launches = None, launches_per_step

    # e2e: host buffers through the C-ABI (pinned host memory), H2D + compute + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        h2d = my_cols * n * 32
        d2h = cosets * C * n * 32
        res_host = torch.empty((cosets, C, n, 4), dtype=torch.int64).pin_memory()
        def step_e2e():
            if my_cols:
                trace[:my_cols].copy_(pinned, non_blocking=True)
            step()
            res_host.copy_(out, non_blocking=True)
        step_e2e(); barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step_e2e()
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item()) / args.steps
        e2e = {"ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}

    if rank == 0:
        muls = algorithmic_muls(cfg, [])
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
                "config": dict(workload="lde_2^%d_x%d_blowup%d" % (cfg["log_n"], C, B), **cfg,
                               l2="working set %.1f GB per step, far above the 126 MB L2; no flush needed" % (
                                   (1 + B) * C * n * 32 / 1e9),
                               parallelism="columns->all_gather->cosets x%d" % world),
                "proof_gen_s": None, "stage_s": {"lde": ms_per_step * 1e-3},
                "gpu_launches": launches_per_step * args.steps, "clocks": clocks}
        if ntt_ms:
            # every k_ntt_pass launch reads and writes each element of its columns once: 64 B per element
            passes = 2 if cfg["log_n"] > 10 else 1
            bytes_total = (1 + B) * C * n * 64.0 * passes
            ach = bytes_total / (ntt_ms * 1e-3) / 1e9
            line["roofline"] = {"bound": "hbm", "kernel": "k_ntt_pass", "achieved": ach, "peak": peak, "unit": "GB/s",
                                "frac": ach / peak, "traffic": None,
                                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650",
                                "launches": n_pass_launches, "avg_launch_ms": ntt_ms / max(1, n_pass_launches),
                                "field_mul_per_s": muls / (ntt_ms * 1e-3)}
        if e2e:
            line["e2e"] = {"value": muls / (e2e["ms_per_step"] * 1e-3), "unit": UNIT,
                           "h2d_bytes_per_step": e2e["h2d_bytes_per_step"] * world if world > 1 else e2e["h2d_bytes_per_step"],
                           "d2h_bytes_per_step": e2e["d2h_bytes_per_step"] * world, "ms_per_step": e2e["ms_per_step"]}
        if world == 1 and not args.no_cpu:
            v, cores, desc, _dt = cpu_sample(cfg)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
