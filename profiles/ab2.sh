#!/usr/bin/env bash
run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$1', round(d['value'],1), round(d['e2e']['value'],1), {k: round(v,2) for k,v in d['config']['stage_ms_per_step'].items()}, round(d['roofline']['avg_launch_ms'],3))"; }
for w in 16 24 32; do
SCS_NO_D2H=1 SCS_EMIT_WARPS=$w SCS_NO_PLAN_SKIP=1 run "noD2H warps=$w redo"
SCS_NO_D2H=1 SCS_EMIT_WARPS=$w run "noD2H warps=$w skip"
done
SCS_EMIT_WARPS=32 run "full warps=32 skip"
