#!/usr/bin/env bash
# same-box A/B: the shipped library against the build of call 11 (before the relay plumbing / 28 warps), 1/20-scale workload, twice
set -uo pipefail
mkdir -p gpurun_out
run() { python bench.py --steps 3 --warmup 3 --scale 0.05 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'value', round(d['value'],1), 'reads_ms', round(d['detail']['stage_ms_per_step_rank0']['reads'],1), 'emit_ms', round(d['roofline']['avg_launch_ms'],4), 'd2h_peak', round(d['roofline']['d2h']['peak'],1), 'frac', round(d['roofline']['d2h']['frac'],3))"; }
for i in 1 2; do
  run shipped
  SCS_LIB_PATH=$PWD/scssim_b200/variants/libscssim_b200_w24.so run call11_w24
  SCS_LIB_PATH=$PWD/scssim_b200/variants/libscssim_b200_w28.so run call11_w28
done
