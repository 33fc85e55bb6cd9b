#!/usr/bin/env bash
# A/B on one box: env-var variants of the same build, kernel-only (SCS_NO_D2H=1) and full
run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$1', round(d['value'],1), round(d['e2e']['value'],1), {k: round(v,2) for k,v in d['config']['stage_ms_per_step'].items()}, round(d['roofline']['avg_launch_ms'],3))"; }
SCS_NO_D2H=1 SCS_NO_PLAN_SKIP=1 run "noD2H redo-indel "
SCS_NO_D2H=1 run "noD2H skip-indel "
SCS_NO_PLAN_SKIP=1 run "full  redo-indel "
run "full  skip-indel "
