#!/usr/bin/env bash
# round 2, final 1-GPU evidence: full-scale bench (configs[3]) with the shipped build, the reference arm, then on the 1/20-scale
# workload the ncu launch list (bounded) and one --set full capture of the emit kernel with source
set -uo pipefail
mkdir -p gpurun_out
( time python bench.py --steps 3 --warmup 3 ) > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench full rc=$?"; tail -4 gpurun_out/r02_bench_n1_final.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --scale 0.05 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/r02_final_small.json 2> gpurun_out/r02_final_small.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/r02_final_launches.csv $CMD > gpurun_out/r02_final_ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'emit_kernel' -s 30 -c 2 -f -o gpurun_out/prof_r02_final $CMD > gpurun_out/r02_final_ncu_full.log 2>&1
echo "full rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1_final.json'))
print(round(d['value'],1), round(d['ms_per_step']), round(d['detail']['fastq_GBps'],2), d['roofline']['d2h']['frac'], d['e2e'], d['roofline']['avg_launch_ms'])
print(json.dumps(d.get('configs1')), json.dumps(d.get('configs1_gz')), json.dumps(d.get('e2e_files'))[:700])
r=json.load(open('gpurun_out/r02_bench_reference.json')); print(r['value'], r['steps'], r['warmup'], r['ms_per_step'], r['cpu_baseline']['stage_s'])
PY
