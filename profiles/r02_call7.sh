#!/usr/bin/env bash
# round 2, call 7: GPU suite on the default build and again on the CHECKED build (every kernel index asserted; compute-sanitizer
# is closed on this pool), gzip output, CTA-size A/B, bench with gzip
set -uo pipefail
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest7.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest7.log
SCS_LIB_PATH=$PWD/scssim_b200/variants/libscssim_b200_checked.so python -m pytest tests -m gpu -x -q -k "not cli and not simuvars" > gpurun_out/r02_pytest7_checked.log 2>&1; echo "checked pytest rc=$?"; tail -4 gpurun_out/r02_pytest7_checked.log
SCS_LIB_PATH=$PWD/scssim_b200/variants/libscssim_b200_checked.so python profiles/sanitize_case.py > gpurun_out/r02_checked_case.log 2>&1; echo "checked case rc=$?"; tail -2 gpurun_out/r02_checked_case.log
bash profiles/ab_warps.sh run 24 28 2>&1 | tee gpurun_out/r02_ab_warps2.txt
python bench.py --steps 2 --warmup 3 --scale 0.05 --no-cpu-baseline > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench7.err
python bench.py --steps 2 --warmup 3 --scale 0.05 --no-cpu-baseline --no-extras --gz > gpurun_out/r02_bench7_gz.json 2> gpurun_out/r02_bench7_gz.err; echo "bench gz rc=$?"; tail -3 gpurun_out/r02_bench7_gz.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench7.json')); print('plain', d['value'], d['roofline']['avg_launch_ms'], json.dumps(d.get('configs1_gz')), json.dumps(d.get('e2e_files'))[:600])
d=json.load(open('gpurun_out/r02_bench7_gz.json')); print('gz', d['value'], d['ms_per_step'], d['detail']['fastq_GBps'], d['detail']['stage_ms_per_step_rank0'])
PY
