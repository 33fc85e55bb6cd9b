#!/usr/bin/env bash
# round 2, call 8 (8 GPUs): host-link probe, bench at N=8/4/2 on the full configs[3] cell, N=8 with gzip output, CLI --gpus 8 over NCCL
set -uo pipefail
mkdir -p gpurun_out
{ nvidia-smi -L; nproc; free -g | head -2; nvidia-smi topo -m | head -12; } > gpurun_out/r02_n8_box.txt 2>&1
timeout 200 ./profiles/d2h_probe > gpurun_out/r02_d2h_probe_n8.json 2> gpurun_out/r02_d2h_probe_n8.err; echo "probe rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
  ( time timeout 400 $TR --nproc-per-node $n --master-port $((29520 + n)) bench.py --gpus $n --steps 2 --warmup 3 --no-extras --no-cpu-baseline ) > gpurun_out/r02_bench_n${n}_full.json 2> gpurun_out/r02_bench_n${n}_full.err
  echo "bench n=$n rc=$?"; tail -4 gpurun_out/r02_bench_n${n}_full.err | cut -c1-300
done
NCCL_DEBUG=INFO timeout 300 $TR --nproc-per-node 8 --master-port 29540 bench.py --gpus 8 --steps 2 --warmup 3 --no-extras --no-cpu-baseline --gz > gpurun_out/r02_bench_n8_gz.json 2> gpurun_out/r02_bench_n8_gz.err; echo "bench n8 gz rc=$?"
grep -E "NCCL INFO (comm 0x.* rank 0 nRanks|NVLS multicast support is available on dev 0|Connected all rings)" gpurun_out/r02_bench_n8_gz.err | head -4
# the drop-in CLI on a 1/16-scale cell (388 Mb FASTA, 77.5 M reads, 24.6 GB of FASTQ): 8 worker threads, NCCL inside the library, one pair of files
PROF=$(python profiles/make_cell_fasta.py 0.0625 /dev/shm/cell.fa | tail -1)
for out in /dev/shm/cli8 /tmp/cli8; do
  rm -f /dev/shm/cell.fa.fai
  ( time NCCL_DEBUG=INFO timeout 300 scssim_b200/bin/scssim genreads -i /dev/shm/cell.fa -m $PROF -r 2e-10 -c 60 -l PE -s 260 -t 8 --seed 7 --gpus 8 -o $out ) > gpurun_out/r02_cli8_$(basename $(dirname $out)).log 2>&1
  echo "cli rc=$? -> $out"; ls -la ${out}_1.fq ${out}_2.fq | awk '{print $5}' | tr '\n' ' '; grep real gpurun_out/r02_cli8_$(basename $(dirname $out)).log
  head -c 400 ${out}_1.fq | head -4 > gpurun_out/r02_cli8_head_$(basename $(dirname $out)).txt; rm -f ${out}_1.fq ${out}_2.fq
done
( time timeout 300 scssim_b200/bin/scssim genreads -i /dev/shm/cell.fa -m $PROF -r 2e-10 -c 60 -l PE -s 260 -t 8 --seed 7 --gpus 8 --gz -o /tmp/cli8gz ) > gpurun_out/r02_cli8_gz.log 2>&1; echo "cli gz rc=$?"; ls -la /tmp/cli8gz* | awk '{s+=$5} END{print s}'; grep real gpurun_out/r02_cli8_gz.log
rm -f /tmp/cli8gz* /dev/shm/cell.fa*
python - <<'PY'
import json
for n in (8,4,2):
    try:
        d=json.load(open(f'gpurun_out/r02_bench_n{n}_full.json')); print(n, round(d['value'],1), round(d['ms_per_step']), round(d['e2e']['value'],1), d['detail']['device_map'].get('order'), d['detail']['share_of_reads_per_rank'], d['roofline']['d2h']['per_gpu_concurrent'], round(d['roofline']['d2h']['frac'],3), round(d['roofline']['d2h']['frac_of_n_x_single_link'],3))
    except Exception as e: print(n,'ERR',e)
try:
    d=json.load(open('gpurun_out/r02_bench_n8_gz.json')); print('gz8', round(d['value'],1), round(d['ms_per_step']), d['detail']['fastq_GBps'], d['detail']['stage_ms_per_step_rank0'])
except Exception as e: print('gz ERR', e)
PY
