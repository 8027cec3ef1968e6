#!/bin/bash
# One GPU-box session: parity tests, probes, bench, ncu launch list, per-kernel pipe metrics and a full capture of the
# dominant kernel.  Usage (from the repo root on the GPU box):  bash tools/gpu_round.sh <tag> [kernel-regex]
# Cost: about 17 GPU-minutes, 11 of them the per-kernel pipe-metric pass (ncu replays all launches of four proofs);
# set SKIP_PIPES=1 to leave that pass out (about 6 minutes).
TAG=${1:-r1}
KREGEX=${2:-k_ntt_pass}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.txt
echo "=== probe"; timeout 300 python tools/gpu_probe.py 2>&1 | tail -40 > gpurun_out/${TAG}_probe.txt; cp gpurun_out/probe.json gpurun_out/${TAG}_probe.json
echo "=== microbench"; tools/build/microbench3 > gpurun_out/${TAG}_microbench3.json 2>&1; tools/build/mulbench > gpurun_out/${TAG}_mulbench.json 2>&1
echo "=== bench"; timeout 900 python bench.py 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json
tail -5 gpurun_out/${TAG}_bench.err
echo "=== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tee gpurun_out/${TAG}_bench_reference.json
echo "=== crypto bench"; timeout 600 python tools/gpu_crypto_bench.py > gpurun_out/${TAG}_crypto.json 2>&1; tail -3 gpurun_out/${TAG}_crypto.json
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bench.log | cut -c1-300
echo "=== ncu pipes (all kernels of one step)"
if [ -z "$SKIP_PIPES" ]; then
  bash tools/ncu_pipes.sh ${TAG} > /dev/null 2>&1
  head -4 gpurun_out/${TAG}_pipes_summary.txt
fi
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 60 -c 2 -f -o gpurun_out/${TAG}_prof \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-200
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
ls -la gpurun_out | head -40
