#!/usr/bin/env bash
# round 2, call 3: the whole GPU suite on the rewritten emit kernel (TMA-staged windows, two-deep slot pipeline), its launch time
# on a 1/20-scale cell, and a --set full capture with source
set -uo pipefail
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest3.log
CMD="python bench.py --steps 2 --warmup 3 --scale 0.05 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err && \
ncu --set full --clock-control none --import-source on -k regex:'emit_kernel' -s 30 -c 2 -f -o gpurun_out/prof_r02_v2 $CMD > gpurun_out/r02_ncu_full3.log 2>&1
echo "ncu rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench3.json'))
print(d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['launches_per_step'], d['detail']['stage_ms_per_step_rank0'])
PY
