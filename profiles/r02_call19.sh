#!/usr/bin/env bash
# round 2, call 19: amplify kernel v2 (8-mer by word loads, GC index, shared Philox in the site search, compile-time replay):
# parity (default + checked build), then configs[2] at 1/10 scale and full scale
set -uo pipefail
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest19.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest19.log
SCS_LIB_PATH=$PWD/scssim_b200/variants/libscssim_b200_checked.so python -m pytest tests -m gpu -x -q -k "pipeline or replay or oracle or distribution" > gpurun_out/r02_pytest19_checked.log 2>&1; echo "checked rc=$?"; tail -3 gpurun_out/r02_pytest19_checked.log
SCS_TRACE=1 python profiles/config3_scaled.py 310000000 1 > gpurun_out/r02_config2_tenth_v2.json 2> gpurun_out/r02_config2_tenth_v2_passes.txt; echo "rc=$?"
cat gpurun_out/r02_config2_tenth_v2.json; grep "scs trace" gpurun_out/r02_config2_tenth_v2_passes.txt | tail -4
