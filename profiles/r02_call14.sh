#!/usr/bin/env bash
# round 2, call 14 (2 GPUs): NVLink relay test, NCCL tests, bench N=2 with a forced relay (rank 1 through rank 0's GPU) for timing
set -uo pipefail
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q > gpurun_out/r02_pytest14.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest14.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR bench.py --gpus 2 --steps 2 --warmup 3 --scale 0.1 --no-extras --no-cpu-baseline > gpurun_out/r02_bench14_n2.json 2> gpurun_out/r02_bench14_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r02_bench14_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench14_n2.json')); print(round(d['value'],1), round(d['ms_per_step']), d['detail']['device_map'], d['detail']['relay'], d['detail']['share_of_reads_per_rank'], d['roofline']['d2h']['per_gpu_concurrent'])
PY
