#!/usr/bin/env bash
# round 2, call 1: box facts, the GPU parity suite, and a --set full capture (with source) of the emit kernel as it stood at
# the start of the round (baseline for the per-line hot table)
set -uo pipefail
mkdir -p gpurun_out
{ nvidia-smi -L; nproc; free -g | head -2; df -h /tmp /dev/shm . 2>/dev/null; mount | grep -E " /tmp| /dev/shm| / " ; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core"; \
  (numactl -H 2>&1 | head -12); grep -i huge /proc/meminfo; nvidia-smi topo -m 2>&1 | head -14; (lspci -tv 2>&1 | head -60); } > gpurun_out/r02_box.txt 2>&1
( dd if=/dev/zero of=/tmp/ddtest bs=16M count=128 oflag=direct 2>&1 | tail -1; dd if=/dev/zero of=/tmp/ddtest bs=16M count=128 2>&1 | tail -1; rm -f /tmp/ddtest; \
  dd if=/dev/zero of=/dev/shm/ddtest bs=16M count=128 2>&1 | tail -1; rm -f /dev/shm/ddtest ) >> gpurun_out/r02_box.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest1.log
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/r02_plain1.log 2> gpurun_out/r02_plain1.err && \
ncu --set full --clock-control none --import-source on -k regex:'emit_kernel' -s 30 -c 2 -f -o gpurun_out/prof_r02_base $CMD > gpurun_out/r02_ncu_full1.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/r02_plain1.log | cut -c1-400
