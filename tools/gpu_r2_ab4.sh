#!/bin/bash
# A/B of the 2^9 compile-time tile for the 2^18 transform (BASELINE configs[1]) against the generic pass kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "ntt or lde or emul or headline" 2>&1 | tail -5
cat > /tmp/cfg1.py <<'PY'
import sys, json, numpy as np, torch
sys.path.insert(0, "tests")
from conftest import rand_felts
from stark_perpetual_b200._lib import get_context
ctx = get_context()
out = {}
for log_n in (16, 18, 20):
    for batch in (1, 64):
        if log_n == 20 and batch == 64: batch = 25
        v = torch.from_numpy(rand_felts(batch << log_n, 1002).view(np.int64)).cuda()
        best = 1e30
        for _ in range(10):
            ctx.ntt_device(v.data_ptr(), log_n, batch)
            best = min(best, ctx.last_kernel_ms)
        out["2^%d x%d" % (log_n, batch)] = round(1e3 * best / batch, 2)
print(json.dumps(out))
PY
for env in "" "SPG_NTT_GENERIC=1" "SPG_NTT_TMA2D=0" "SPG_NTT_TMA2D_STORE=0"; do
  echo "=== $env  (us per vector)"
  env $env python /tmp/cfg1.py 2>&1 | tail -1 | tee -a gpurun_out/r2p_cfg1.txt
done
