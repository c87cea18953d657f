#!/bin/bash
# usage (under gpurun): bash tools/gpu_ncu_full.sh <tag> <demangled-name regex> [skip] [count]
# one `ncu --set full` capture of the matching kernel(s) of a bench step (after a clean plain run)
set -u
tag=$1; re=$2; skip=${3:-4}; cnt=${4:-1}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$re" -s $skip -c $cnt \
  -f -o gpurun_out/full_${tag} python bench.py --quick --steps 1 --warmup 3 > gpurun_out/ncu_full_${tag}.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/full_${tag}.ncu-rep
