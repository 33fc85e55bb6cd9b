#!/usr/bin/env bash
# Run on the GPU box via gpurun: plain bench first (must exit 0), then the ncu launch list and one
# --set full capture of the hot kernels with the SAME command line (B200_PROFILING.md recipe).
set -uo pipefail
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2> gpurun_out/plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:'emit_kernel|plan_kernel|amplify_kernel' -s 30 -c 8 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -2 gpurun_out/plain.log
