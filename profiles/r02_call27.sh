#!/usr/bin/env bash
# round 2, call 27: BASELINE configs[4] share of one GPU (simuvars -> genreads PE100 60x, XTen profile) and the final full-scale bench
set -uo pipefail
mkdir -p gpurun_out
timeout 300 python profiles/config4_pipeline.py > gpurun_out/r02_config4_pipeline_n1.json 2> gpurun_out/r02_config4_pipeline_n1.err; echo "config4 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_config4_pipeline_n1.json')); p=d['pass1']; print({k: p[k] for k in ('simuvars_to_genome_s','amplify_alloc_s','reads_s','total_s','M_reads_per_s_this_rank','device_ms')})"
( time python bench.py --steps 3 --warmup 3 ) > gpurun_out/r02_bench_n1_final3.json 2> gpurun_out/r02_bench_n1_final3.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1_final3.json')); print(round(d['value'],1), round(d['ms_per_step']), d['roofline']['d2h']['frac'], d['detail']['stage_ms_per_step_rank0'], d['e2e']['value'], d['roofline']['avg_launch_ms'], d['configs1']['value'], d['configs1_gz']['value'], d['cpu_baseline']['value'])"
