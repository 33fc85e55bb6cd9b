#!/usr/bin/env bash
# A/B of the emit kernel's CTA size: builds (here, no GPU needed) one library per warp count into scssim_b200/variants/, then
# (on the GPU box) runs the 1/20-scale bench with each and prints the average emit launch time.
# usage: profiles/ab_warps.sh build 16 20 24 28 32   |   profiles/ab_warps.sh run 16 20 24 28 32
set -uo pipefail
mode=$1; shift
cd "$(dirname "$0")/.."
if [ "$mode" = build ]; then
  for w in "$@"; do
    make -s -C scssim_b200/csrc variant NAME=w$w DEFS=-DSCS_EMIT_WARPS=$w || exit 1
    grep -A2 "emit_kernelILb0ELb0" scssim_b200/variants/reads.w$w.log | grep -E "registers|spill" | tr '\n' ' '; echo " <- w=$w"
  done
else
  for w in "$@"; do
    SCS_LIB_PATH=$PWD/scssim_b200/variants/libscssim_b200_w$w.so python bench.py --steps 2 --warmup 2 --scale 0.05 --no-extras --no-cpu-baseline 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('w=$w', 'emit_ms', round(d['roofline']['avg_launch_ms'],4), 'reads_ms', round(d['detail']['stage_ms_per_step_rank0']['reads'],1), 'value', round(d['value'],1))"
  done
fi
