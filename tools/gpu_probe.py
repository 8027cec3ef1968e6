#!/usr/bin/env python3
"""Throughput probes on the GPU box: field multiplication rate, IMAD.WIDE rate, NTT timings."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import stark_perpetual_b200 as spg  # noqa: E402
from stark_perpetual_b200._lib import NTT_NAT_TO_REV  # noqa: E402
from conftest import rand_felts  # noqa: E402


def main():
    import torch
    ctx = spg.get_context(0)
    out = {}
    for ch in (1, 2, 4):
        m, w = ctx.bench_field_mul(4000, ch)
        out["mul_per_s_chains%d" % ch] = m
        out["imad_wide_per_s"] = w
    for log_n, batch in ((18, 1), (18, 64), (20, 25), (22, 8)):
        n = 1 << log_n
        x = torch.from_numpy(rand_felts(batch * n, 1).view(np.int64)).cuda()
        ctx.ntt_device(x.data_ptr(), log_n, batch)
        best = 1e9
        for _ in range(5):
            ctx.ntt_device(x.data_ptr(), log_n, batch)
            best = min(best, ctx.last_kernel_ms)
        muls = batch * (n // 2) * log_n
        out["ntt_2^%d_x%d" % (log_n, batch)] = {
            "ms": best, "field_mul_per_s": muls / (best * 1e-3), "GBps": batch * n * 64 / (best * 1e-3) / 1e9}
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
