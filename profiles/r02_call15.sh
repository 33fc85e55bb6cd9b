#!/usr/bin/env bash
# round 2, call 15 (8 GPUs): bench at N=8 (GPUs behind the slow host links relay over NVLink through the fast group), N=4 and N=2 on
# the best-connected GPUs; full configs[3] cell
set -uo pipefail
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
  ( time timeout 400 $TR --nproc-per-node $n --master-port $((29620 + n)) bench.py --gpus $n --steps 2 --warmup 4 --no-extras --no-cpu-baseline ) > gpurun_out/r02_relay_bench_n${n}.json 2> gpurun_out/r02_relay_bench_n${n}.err
  echo "bench n=$n rc=$?"; tail -3 gpurun_out/r02_relay_bench_n${n}.err | cut -c1-200
done
python - <<'PY'
import json
for n in (8,4,2):
    try:
        d=json.load(open(f'gpurun_out/r02_relay_bench_n{n}.json')); r=d['roofline']['d2h']
        print(n, round(d['value'],1), round(d['ms_per_step']), 'e2e', round(d['e2e']['value'],1), d['detail']['device_map'].get('order'), d['detail']['relay'], d['detail']['share_of_reads_per_rank'], r['per_gpu_concurrent'], round(r['achieved'],1), r['peak_of_links_used_with_relay'])
    except Exception as e: print(n,'ERR',e)
PY
