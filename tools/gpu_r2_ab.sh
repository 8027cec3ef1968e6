for v in "" _v1 _v2; do
  echo "=== aux with libspg$v.so"
  SPG_LIB=$PWD/stark_perpetual_b200/libspg$v.so timeout 600 python - <<'PY' 2>&1 | tail -12
import json, os, sys
sys.path.insert(0, 'tools'); sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import aux_bench, stark_perpetual_b200 as spg
a = aux_bench.measure(spg.get_context(0), with_reference=False)
for k in ("cfg0_pedersen_1024", "pedersen_2^20", "cfg4_orders_valid_mix", "cfg4_orders_invalid_mix"):
    print(k, {x: (round(y, 3) if isinstance(y, float) else y) for x, y in a[k].items() if x in ("ms", "e2e_ms", "hash_per_s", "orders_per_s", "statuses_as_expected", "oracle_sample_ok", "bad_status")})
json.dump(a, open("gpurun_out/r2f_aux%s.json" % os.environ["SPG_LIB"].split("libspg")[-1].replace(".so", ""), "w"))
PY
done
echo "=== ncu full on 25-column launches"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ntt_tile -s 74 -c 3 -f -o gpurun_out/r2f_prof python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-aux --no-verify > gpurun_out/r2f_ncu_full.log 2>&1
ncu -i gpurun_out/r2f_prof.ncu-rep --page raw --csv > gpurun_out/r2f_prof_raw.csv 2>/dev/null
ncu -i gpurun_out/r2f_prof.ncu-rep --page source --csv --print-source sass > gpurun_out/r2f_prof_src.csv 2>/dev/null
python tools/ncu_stalls.py gpurun_out/r2f_prof_src.csv | awk 'NR<=34' | tee gpurun_out/r2f_stalls.txt | head -20
echo "=== new tests"; timeout 600 python -m pytest tests/test_cairo_artifacts.py tests/test_gpu_state_tree.py -m gpu -q 2>&1 | tail -4
