#!/usr/bin/env bash
# round 2, call 22: free-running indel stage by distance to the next event (oracle + kernel): parity, distribution vs the
# reference, checked build, then emit-kernel time on the bench workload (1/20 scale) and the gzip extra
set -uo pipefail
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest22.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest22.log
SCS_LIB_PATH=$PWD/scssim_b200/variants/libscssim_b200_checked.so python -m pytest tests -m gpu -x -q -k "pipeline or replay or distribution" > gpurun_out/r02_pytest22_checked.log 2>&1; echo "checked rc=$?"; tail -3 gpurun_out/r02_pytest22_checked.log
python bench.py --steps 3 --warmup 3 --scale 0.05 --no-cpu-baseline > gpurun_out/r02_bench22_scale005.json 2> gpurun_out/r02_bench22.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench22_scale005.json").read())
print("value", round(d["value"], 1), "emit_ms", d["roofline"].get("avg_launch_ms"), "stages", d["detail"]["stage_ms_per_step_rank0"])
for k in ("configs1", "configs1_gz"):
    if k in d.get("extras", {}): print(k, {a: d["extras"][k].get(a) for a in ("value", "fastq_GBps", "ms_per_step")})
PY
