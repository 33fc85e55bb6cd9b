#!/usr/bin/env bash
# round 2, call 18: configs[2] at 1/10 scale — pass timings, then one ncu --set full capture of the last two amplify passes
set -uo pipefail
mkdir -p gpurun_out
SCS_TRACE=1 python profiles/config3_scaled.py 310000000 1 > gpurun_out/r02_config2_tenth.json 2> gpurun_out/r02_config2_tenth_passes.txt; echo "rc=$?"
cat gpurun_out/r02_config2_tenth.json; grep "scs trace" gpurun_out/r02_config2_tenth_passes.txt | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:amplify_kernel -s 8 -c 2 -f -o gpurun_out/prof_r02_amplify \
  python profiles/config3_scaled.py 310000000 1 > gpurun_out/ncu_amplify.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_amplify.log
ls -la gpurun_out/prof_r02_amplify.ncu-rep
