#!/usr/bin/env bash
# round 2, call 14 (2 GPUs): NVLink relay test + NCCL CLI tests, then bench N=2 with rank 1's output forced through rank 0's GPU
set -uo pipefail
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q > gpurun_out/r02_pytest14.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest14.log
