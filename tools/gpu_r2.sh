#!/bin/bash
# Round-2 GPU session: parity tests, bench (own + reference arm), ncu launch list and one full capture of the dominant
# kernel.  Usage (repo root on the GPU box):  bash tools/gpu_r2.sh <tag> [kernel-regex]
TAG=${1:-r2}
KREGEX=${2:-k_ntt_pass}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc >> gpurun_out/${TAG}_smi.txt
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -40 | tee gpurun_out/${TAG}_pytest.txt
echo "=== bench"; timeout 900 python bench.py 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | cut -c1-1500
tail -5 gpurun_out/${TAG}_bench.err
if [ -z "$SKIP_REF" ]; then
  echo "=== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tee gpurun_out/${TAG}_bench_reference.json | cut -c1-600
fi
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-aux --no-verify > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bench.log | cut -c1-300
python tools/ncu_summarize.py gpurun_out/${TAG}_launches.csv 2>/dev/null | head -30 | tee gpurun_out/${TAG}_launch_summary.txt
if [ -z "$SKIP_FULL" ]; then
  echo "=== ncu full"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 60 -c 3 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-aux --no-verify > gpurun_out/${TAG}_ncu_full.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-200
  ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
fi
ls -la gpurun_out | head -30
