#!/usr/bin/env bash
# round 2, call 24: GPU suite on the final tree (persistent allocation-stage scratch), checked build subset, stage times of the bench workload
set -uo pipefail
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest24.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest24.log
SCS_LIB_PATH=$PWD/scssim_b200/variants/libscssim_b200_checked.so python -m pytest tests -m gpu -x -q -k "pipeline or multirank" > gpurun_out/r02_pytest24_checked.log 2>&1; echo "checked rc=$?"; tail -3 gpurun_out/r02_pytest24_checked.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --scale 0.25 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'stages', d['detail']['stage_ms_per_step_rank0'], 'emit_ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['d2h']['frac'],3))"
