#!/usr/bin/env bash
# round 2, call 10: gzip tests on the faster deflate kernel, its timing, and a dry run of the CLI recipe used on the 8-GPU box
set -uo pipefail
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "gzip or gz_flag" > gpurun_out/r02_pytest10.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest10.log
python bench.py --steps 2 --warmup 3 --scale 0.05 --no-cpu-baseline --no-extras --gz > gpurun_out/r02_bench10_gz.json 2> gpurun_out/r02_bench10_gz.err; echo "bench gz rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench10_gz.json')); print('gz', d['value'], d['ms_per_step'], d['detail']['fastq_GBps'], d['detail']['stage_ms_per_step_rank0'], d['roofline']['avg_launch_ms'])
PY
CMD="python bench.py --steps 1 --warmup 1 --scale 0.01 --no-cpu-baseline --no-extras --gz"
$CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'deflate|compact' -s 20 -c 40 --csv --log-file gpurun_out/r02_gz_launches2.csv $CMD > /dev/null 2>&1
grep -E "deflate|compact" gpurun_out/r02_gz_launches2.csv | awk -F'","' '{print $5, $(NF)}' | sed 's/"//g' | awk '{n[$1]++; s[$1]+=$NF} END{for(k in n) print k, n[k], s[k]/n[k]}'
PROF=$(python profiles/make_cell_fasta.py 0.002 /dev/shm/cell.fa | tail -1); echo "profile=$PROF"; ls -la /dev/shm/cell.fa
( time scssim_b200/bin/scssim genreads -i /dev/shm/cell.fa -m $PROF -r 2e-10 -c 60 -l PE -s 260 -t 8 --seed 7 -o /dev/shm/cli1 ) 2>&1 | tail -6; ls -la /dev/shm/cli1_*; rm -f /dev/shm/cli1_* /dev/shm/cell.fa*
