#!/usr/bin/env bash
# round 2, call 6: GPU suite on the current tree (subst pass, geometric error skipping), then compute-sanitizer memcheck on the smallest case
set -uo pipefail
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest6.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest6.log
python profiles/sanitize_case.py > gpurun_out/r02_sanitize_plain.log 2>&1 && \
timeout 900 compute-sanitizer --tool memcheck --leak-check no --log-file gpurun_out/r02_sanitizer_memcheck.txt python profiles/sanitize_case.py > gpurun_out/r02_sanitize_memcheck_run.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/r02_sanitize_memcheck_run.log; tail -5 gpurun_out/r02_sanitizer_memcheck.txt
bash profiles/ab_warps.sh run 24 28 2>&1 | tee gpurun_out/r02_ab_warps2.txt
