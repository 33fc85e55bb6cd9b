#!/usr/bin/env bash
# round 2, call 11: GPU suite on the current tree (4-base decode, caps 320/480/64, RL 300 case), checked build, warp A/B, bench
set -uo pipefail
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest11.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest11.log
SCS_LIB_PATH=$PWD/scssim_b200/variants/libscssim_b200_checked.so python -m pytest tests -m gpu -x -q -k "pipeline or replay" > gpurun_out/r02_pytest11_checked.log 2>&1; echo "checked pytest rc=$?"; tail -3 gpurun_out/r02_pytest11_checked.log
bash profiles/ab_warps.sh run 24 28 2>&1 | tee gpurun_out/r02_ab_warps3.txt
