#!/usr/bin/env bash
# round 2, call 4: GPU suite (new: file sink, one-file multi-rank output, sizing pass), CTA-size A/B of the emit kernel, ncu capture
set -uo pipefail
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest4.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest4.log
bash profiles/ab_warps.sh run 16 20 24 28 32 2>&1 | tee gpurun_out/r02_ab_warps.txt
CMD="python bench.py --steps 2 --warmup 3 --scale 0.05 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err && \
ncu --set full --clock-control none --import-source on -k regex:'emit_kernel' -s 30 -c 2 -f -o gpurun_out/prof_r02_v3 $CMD > gpurun_out/r02_ncu_full4.log 2>&1
echo "ncu rc=$?"
