#!/bin/bash
# second AIR (ECDSA builtin): parity tests, then trace / proof timings at 2^16 .. 2^20
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ecdsa_air.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python - <<'PY' 2>&1 | tail -12
import json, sys
sys.path.insert(0, 'tools'); sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import aux_bench, stark_perpetual_b200 as spg
ctx = spg.get_context(0)
out = {}
for log_n in (16, 20):
    row = aux_bench.ecdsa_air(ctx, log_n)
    out["2^%d" % log_n] = row
    print(json.dumps(row))
json.dump(out, open("gpurun_out/r2q_ecdsa_air.json", "w"), indent=1)
PY
