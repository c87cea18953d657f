#!/bin/bash
# usage (under gpurun): bash tools/gpu_profile.sh <tag>
# plain run first (must exit 0), then the ncu launch list of the same command (B200_PROFILING.md recipe)
set -u
tag=${1:-rXX}
mkdir -p gpurun_out
python bench.py --quick --steps 1 --warmup 3 > gpurun_out/plain_${tag}.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_${tag}.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_b18944_${tag}.csv \
  python bench.py --quick --steps 1 --warmup 3 > gpurun_out/ncu_${tag}.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_b18944_${tag}.csv
