#!/bin/bash
# Per-kernel pipe utilisation of one bench step (all kernels): ncu raw metrics -> CSV in gpurun_out/<tag>_pipes.csv
TAG=${1:-r1}
mkdir -p gpurun_out
M="gpu__time_duration.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"
# every launch of the process is captured; the summary keeps the last proof (bench prints its launch count)
timeout 1800 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${TAG}_pipes.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-verify > gpurun_out/${TAG}_pipes.log 2>&1
tail -2 gpurun_out/${TAG}_pipes.log | cut -c1-200
LAST=$(tail -1 gpurun_out/${TAG}_pipes.log | python -c "import json,sys; print(json.loads(sys.stdin.read())['gpu_launches'])")
python tools/ncu_summarize.py gpurun_out/${TAG}_pipes.csv --last ${LAST} > gpurun_out/${TAG}_pipes_summary.txt 2>&1
ls -la gpurun_out/${TAG}_pipes.csv
