#!/usr/bin/env bash
# round 2, call 21: new lane/warp mixed-pass test; configs[2] at full scale with the new amplification kernels; ncu of the lane kernel
set -uo pipefail
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "primers" > gpurun_out/r02_pytest21.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest21.log
SCS_TRACE=1 timeout 600 python profiles/config3_scaled.py 1550000000 2 > gpurun_out/r02_config2_full_v3.json 2> gpurun_out/r02_config2_full_v3_passes.txt; echo "config2 rc=$?"
cat gpurun_out/r02_config2_full_v3.json; grep "scs trace" gpurun_out/r02_config2_full_v3_passes.txt | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:amplify_semis_lanes -s 4 -c 1 -f -o gpurun_out/prof_r02_lanes \
  python profiles/config3_scaled.py 310000000 1 > gpurun_out/ncu_lanes.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_lanes.log
