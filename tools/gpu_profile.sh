#!/bin/bash
# usage (under gpurun): bash tools/gpu_profile.sh <tag> [batch] [0 = launch list | 1 = launch list + GEMM metrics | 2 = GEMM metrics only]
# (one profiler pass per gpurun call: run mode 0 and mode 2 as two calls)
# plain run first (must exit 0), then the ncu launch list of the same command (B200_PROFILING.md recipe) and, optionally,
# DRAM bytes / duration / tensor-pipe activity of the GEMM launches of one step.  Also records, for the summariser:
#   gpurun_out/shapes_<tag>.csv     one step's per-launch (family, label, M, N, K, us, work) from the library's own event timing
#   gpurun_out/csrc_hash_<tag>.txt  hash of the CUDA sources the capture ran on (bench.py refuses a stale ncu_traffic.json)
set -u
tag=${1:-rXX}; batch=${2:-37888}; gm=${3:-0}
mkdir -p gpurun_out
cmd="python bench.py --quick --steps 1 --warmup 3 --batch $batch"
rm -f gpurun_out/shapes_${tag}.csv
MFAC_PROFILE_CSV=gpurun_out/shapes_${tag}.csv $cmd > gpurun_out/plain_${tag}.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_${tag}.log; exit 1; }
python -c "import bench; print(bench.csrc_hash())" > gpurun_out/csrc_hash_${tag}.txt
ngemm=$(grep -c '^0,' gpurun_out/shapes_${tag}.csv)
echo "GEMM launches per step: $ngemm"
if [ "$gm" != 2 ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_b${batch}_${tag}.csv \
  $cmd > gpurun_out/ncu_${tag}.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_b${batch}_${tag}.csv
fi
if [ "$gm" = 1 ] || [ "$gm" = 2 ]; then
  # skip the warm-up steps' GEMMs (3 warm-up + 1 timed step precede the profiled one? no: -s counts matching launches; the
  # e2e leg runs 2 + 1 more steps) -- take the LAST full step by skipping (steps_before) * ngemm launches
  skip=$((ngemm * 2))
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:gemm_tcgen05 -s $skip -c $ngemm --csv --log-file gpurun_out/gemm_step_metrics_b${batch}_${tag}.csv \
    $cmd > gpurun_out/ncu_gm_${tag}.log 2>&1
  echo "ncu gemm metrics rc=$?"; wc -l gpurun_out/gemm_step_metrics_b${batch}_${tag}.csv
fi
