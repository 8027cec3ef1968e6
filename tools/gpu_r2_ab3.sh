for v in "" _old "" _old; do
  echo "=== bench with libspg$v.so"
  SPG_LIB=$PWD/stark_perpetual_b200/libspg$v.so timeout 600 python bench.py --no-cpu --no-aux --no-verify --no-e2e 2>/dev/null | tee gpurun_out/r2l_bench$v.json | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['ms_per_step'], 3), {k: round(v, 2) for k, v in d['stage_ms'].items()})"
done
echo "=== SPG_NTT_TMA2D=0 SPG_NTT_TMA=0 (per-thread loads everywhere)"
SPG_NTT_TMA2D=0 SPG_NTT_TMA=0 timeout 600 python bench.py --no-cpu --no-aux --no-verify --no-e2e 2>/dev/null | tee gpurun_out/r2l_bench_notma.json | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['ms_per_step'], 3), {k: round(v, 2) for k, v in d['stage_ms'].items()})"
