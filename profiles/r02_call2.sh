#!/usr/bin/env bash
# round 2, call 2: the rest of the GPU suite, then the reworked bench (configs[3]) first at 1/50 scale, then at full scale
set -uo pipefail
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_simuvars.py -k "resampled or slab_limits or properties or replay or multirank or units or cli or distribution" > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest2.log
python bench.py --steps 2 --warmup 3 --scale 0.02 > gpurun_out/r02_bench_small.json 2> gpurun_out/r02_bench_small.err; echo "bench small rc=$?"; tail -3 gpurun_out/r02_bench_small.err; cut -c1-600 gpurun_out/r02_bench_small.json
( time python bench.py --steps 2 --warmup 3 ) > gpurun_out/r02_bench_full_n1.json 2> gpurun_out/r02_bench_full_n1.err; echo "bench full rc=$?"; tail -5 gpurun_out/r02_bench_full_n1.err; cut -c1-300 gpurun_out/r02_bench_full_n1.json
