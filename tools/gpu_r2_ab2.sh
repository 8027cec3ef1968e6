for v in "" _v3 _v4; do
  echo "=== bench with libspg$v.so"
  SPG_LIB=$PWD/stark_perpetual_b200/libspg$v.so timeout 600 python bench.py --no-cpu --no-aux --no-verify --no-e2e 2>/dev/null | tee gpurun_out/r2i_bench$v.json | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['ms_per_step'], 3), {k: round(v, 2) for k, v in d['stage_ms'].items()})"
done
