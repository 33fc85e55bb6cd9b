#!/usr/bin/env bash
# round 2, call 29: A/B of plan_slot's batched single draws (variant `batch`) against the default build: parity of both
# (pipeline / multirank / properties tests), emit-kernel time on the bench workload at 1/20 scale, twice each
set -uo pipefail
mkdir -p gpurun_out
V=$PWD/scssim_b200/variants/libscssim_b200_batch.so
python -m pytest tests -m gpu -x -q -k "pipeline or multirank or properties or units" > gpurun_out/r02_pytest29_default.log 2>&1; echo "default rc=$?"; tail -2 gpurun_out/r02_pytest29_default.log
SCS_LIB_PATH=$V python -m pytest tests -m gpu -x -q -k "pipeline or multirank or properties or units or distribution" > gpurun_out/r02_pytest29_batch.log 2>&1; echo "batch rc=$?"; tail -2 gpurun_out/r02_pytest29_batch.log
run() { python bench.py --steps 3 --warmup 3 --scale 0.05 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'value', round(d['value'],1), 'emit_ms', round(d['roofline']['avg_launch_ms'],4))"; }
for i in 1 2; do run default; SCS_LIB_PATH=$V run batch; done 2>&1 | tee gpurun_out/r02_ab_batch_draws.txt
