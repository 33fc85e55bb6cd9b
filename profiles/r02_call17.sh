#!/usr/bin/env bash
# round 2, call 17: full GPU suite + smoke on the final tree; BASELINE configs[2] (3.1 Gb haploid, default gamma, SE) at full scale
set -uo pipefail
mkdir -p gpurun_out
make -C scssim_b200/csrc -q 2>/dev/null || true
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest17.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest17.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke17.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke17.log
SCS_TRACE=1 timeout 600 python profiles/config3_scaled.py 1550000000 2 > gpurun_out/r02_config2_full.json 2> gpurun_out/r02_config2_full_passes.txt; echo "config2 rc=$?"
cat gpurun_out/r02_config2_full.json; grep "scs trace" gpurun_out/r02_config2_full_passes.txt | tail -12
