#!/usr/bin/env bash
for mb in 32 64 128 512; do
  python bench.py --steps 4 --warmup 2 --no-cpu-baseline --slab-mb $mb 2>/dev/null > /tmp/o.json
  python -c "import json; d=json.load(open('/tmp/o.json')); print('slab_mb=$mb', round(d['value'],1), round(d['e2e']['value'],1), d['config']['stage_ms_per_step'], d['roofline']['avg_launch_ms'])"
done
