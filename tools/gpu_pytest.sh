#!/bin/bash
# usage (under gpurun): bash tools/gpu_pytest.sh [pytest args]
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
