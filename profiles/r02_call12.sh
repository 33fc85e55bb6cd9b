#!/usr/bin/env bash
# debugging the long-read (tables in L2) path
set -uo pipefail
mkdir -p gpurun_out
export SCS_LIB_PATH=$PWD/scssim_b200/variants/libscssim_b200_checked.so CUDA_LAUNCH_BLOCKING=1
echo "--- tables staged in smem, sampled from L2"; SCS_DIAG_SMEM=2 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -k "small" 2>&1 | grep -E "passed|failed|FAILED|error" | head -5
echo "--- tables in L2, no smem tables"; SCS_DIAG_SMEM=0 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -k "small" 2>&1 | grep -E "passed|failed|FAILED|error" | head -5
