#!/usr/bin/env bash
# round 2, call 28: occupancy A/B of the lane kernel (4 CTAs/SM at 64 registers = shipped, 5 at 48, 6 at 40) on configs[2] at 1/10 scale
set -uo pipefail
mkdir -p gpurun_out
for v in shipped lanes_c5 lanes_c6 shipped; do
  if [ $v = shipped ]; then unset SCS_LIB_PATH; else export SCS_LIB_PATH=$PWD/scssim_b200/variants/libscssim_b200_$v.so; fi
  SCS_TRACE=1 python profiles/config3_scaled.py 310000000 1 2> gpurun_out/ab_occ_$v.txt | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'amplify_ms', round(d['ms']['amplify'],1), 'fulls', d['fulls'], 'primers_left', d['primers_left'])"
  grep "from semis" gpurun_out/ab_occ_$v.txt | tail -3 | sed -e 's/.*round/  round/' -e 's/templates.*| scan+alloc [0-9.]* / /'
done 2>&1 | tee gpurun_out/r02_ab_lanes_occupancy.txt
