#!/usr/bin/env bash
# round 2, call 23: (1) A/B of the lane kernel's product-step threshold (1/4, 1/2 = shipped, 3/4, all) on configs[2] at 1/10
# scale; (2) full-scale bench (configs[3]) with the final kernels; (3) ncu launch list (bounded) and one --set full capture of
# the emit kernel on the 1/20-scale bench workload
set -uo pipefail
mkdir -p gpurun_out
for v in shipped lanes_q lanes_t lanes_a; do
  if [ $v = shipped ]; then unset SCS_LIB_PATH; else export SCS_LIB_PATH=$PWD/scssim_b200/variants/libscssim_b200_$v.so; fi
  SCS_TRACE=1 python profiles/config3_scaled.py 310000000 1 2> gpurun_out/ab_lanes_$v.txt | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'amplify_ms', round(d['ms']['amplify'],1), 'fulls', d['fulls'])"
  grep "from semis" gpurun_out/ab_lanes_$v.txt | tail -2 | sed -e 's/.*round/  round/' -e 's/templates.*| scan+alloc [0-9.]* / /'
done 2>&1 | tee gpurun_out/r02_ab_lanes.txt
unset SCS_LIB_PATH
( time python bench.py --steps 3 --warmup 3 ) > gpurun_out/r02_bench_n1_final2.json 2> gpurun_out/r02_bench_n1_final2.err; echo "bench full rc=$?"; tail -4 gpurun_out/r02_bench_n1_final2.err
CMD="python bench.py --steps 2 --warmup 3 --scale 0.05 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/r02_final2_small.json 2> gpurun_out/r02_final2_small.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/r02_final2_launches.csv $CMD > gpurun_out/r02_final2_ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'emit_kernel' -s 30 -c 2 -f -o gpurun_out/prof_r02_final2 $CMD > gpurun_out/r02_final2_ncu_full.log 2>&1
echo "full rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1_final2.json'))
print(round(d['value'],1), round(d['ms_per_step']), round(d['detail']['fastq_GBps'],2), d['roofline']['d2h']['frac'], d['e2e'], d['roofline']['avg_launch_ms'], d['detail']['stage_ms_per_step_rank0'])
print(json.dumps(d.get('configs1')), json.dumps(d.get('configs1_gz')), json.dumps(d.get('e2e_files'))[:700])
PY
