#!/bin/bash
# usage (under gpurun): bash tools/gpu_profile.sh <tag> [batch] [gemm-metrics: 0|1]
# plain run first (must exit 0), then the ncu launch list of the same command (B200_PROFILING.md recipe) and, optionally,
# DRAM bytes / duration / tensor-pipe activity of the 129 GEMM launches of one step
set -u
tag=${1:-rXX}; batch=${2:-37888}; gm=${3:-0}
mkdir -p gpurun_out
cmd="python bench.py --quick --steps 1 --warmup 3 --batch $batch"
$cmd > gpurun_out/plain_${tag}.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_${tag}.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_b${batch}_${tag}.csv \
  $cmd > gpurun_out/ncu_${tag}.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_b${batch}_${tag}.csv
if [ "$gm" = 1 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:gemm_tcgen05 -s 258 -c 129 --csv --log-file gpurun_out/gemm_step_metrics_b${batch}_${tag}.csv \
    $cmd > gpurun_out/ncu_gm_${tag}.log 2>&1
  echo "ncu gemm metrics rc=$?"; wc -l gpurun_out/gemm_step_metrics_b${batch}_${tag}.csv
fi
