#!/bin/bash
# A/B of the deferred-check Pedersen absorption (libspg_v2.so, -DPEDERSEN_STREAM=2) against the set-bit stream (libspg.so)
mkdir -p gpurun_out
for v in "" ${VARIANT:-_v2}; do
  echo "=== libspg$v.so"
  SPG_LIB=$PWD/stark_perpetual_b200/libspg$v.so timeout 300 python -m pytest tests/test_gpu_pedersen.py tests/test_gpu_ecdsa.py -m gpu -x -q 2>&1 | tail -1
  SPG_LIB=$PWD/stark_perpetual_b200/libspg$v.so timeout 600 python - <<'PY' 2>&1 | tail -6
import json, os, sys
sys.path.insert(0, 'tools'); sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import aux_bench, stark_perpetual_b200 as spg
a = aux_bench.measure(spg.get_context(0), with_reference=False)
for k in ("cfg0_pedersen_1024", "pedersen_2^20", "cfg4_orders_valid_mix", "cfg4_orders_invalid_mix"):
    print(k, {x: (round(y, 3) if isinstance(y, float) else y) for x, y in a[k].items() if x in ("ms", "e2e_ms", "statuses_as_expected", "oracle_sample_ok", "bad_status")})
json.dump(a, open("gpurun_out/r2ab6_aux%s.json" % os.environ["SPG_LIB"].split("libspg")[-1].replace(".so", ""), "w"))
PY
done
