#!/usr/bin/env bash
# round 2, call 5 (2 GPUs): NCCL inside the library — CLI --gpus 2, bench at N=2 (small, then full scale), D2H probe
set -uo pipefail
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_n2_gpus.txt
python -m pytest tests/test_gpu_multigpu.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/r02_pytest5.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest5.log
./profiles/d2h_probe > gpurun_out/r02_d2h_probe_n2.json 2> gpurun_out/r02_d2h_probe_n2.err; echo "probe rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
NCCL_DEBUG=INFO $TR bench.py --gpus 2 --steps 2 --warmup 3 --scale 0.05 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_n2_small.json 2> gpurun_out/r02_bench_n2_small.err; echo "bench n2 small rc=$?"
grep -E "NCCL INFO (Connected|comm 0x|ncclCommInitRank|Channel 00/0 :|NVLS)" gpurun_out/r02_bench_n2_small.err | head -12
cut -c1-250 gpurun_out/r02_bench_n2_small.json
( time $TR bench.py --gpus 2 --steps 2 --warmup 3 --no-extras ) > gpurun_out/r02_bench_n2_full.json 2> gpurun_out/r02_bench_n2_full.err; echo "bench n2 full rc=$?"; tail -4 gpurun_out/r02_bench_n2_full.err
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench_n2_small.json','gpurun_out/r02_bench_n2_full.json'):
    try:
        d=json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['detail']['share_of_reads_per_rank'], d['detail']['device_map'], d['roofline']['d2h'])
    except Exception as e: print(f, 'ERR', e)
PY
