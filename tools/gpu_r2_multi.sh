#!/bin/bash
# Multi-GPU session (gpurun --gpus N): byte-identity of the C++ sharded prover, bench at N ranks.
# Usage: bash tools/gpu_r2_multi.sh <tag> <N> [pytest-args]
TAG=${1:-r2m}
N=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ -n "$3" ]; then
  echo "=== pytest $3"; timeout 1200 python -m pytest $3 -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.txt
fi
echo "=== multi_gpu_check 2^14"; timeout 600 $TR --master-port 29511 tools/multi_gpu_check.py 14 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -8 | tee gpurun_out/${TAG}_check14.txt
echo "=== multi_gpu_check 2^20"; timeout 600 $TR --master-port 29512 tools/multi_gpu_check.py 20 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -8 | tee gpurun_out/${TAG}_check20.txt
echo "=== bench --gpus $N"; timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench_n$N.json | cut -c1-900
tail -5 gpurun_out/${TAG}_bench.err
