#!/bin/bash
# ncu --set full on the 25-column coset-transform launches of the final kernels (TMA load + store), then compute-sanitizer
mkdir -p gpurun_out
TAG=${1:-r2u}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ntt_tile -s 112 -c 4 -f -o gpurun_out/${TAG}_prof \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-aux --no-verify > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-200
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_prof_src.csv 2>/dev/null
python tools/ncu_stalls.py gpurun_out/${TAG}_prof_src.csv 2>/dev/null | awk 'NR<=34' > gpurun_out/${TAG}_stalls.txt
rm -f gpurun_out/${TAG}_prof.ncu-rep
echo "=== memcheck"; timeout 1200 compute-sanitizer --tool memcheck python tools/sanitize.py 2>&1 | tail -6 | tee gpurun_out/${TAG}_sanitize_memcheck.txt
echo "=== racecheck"; timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize.py 2>&1 | tail -6 | tee gpurun_out/${TAG}_sanitize_racecheck.txt
