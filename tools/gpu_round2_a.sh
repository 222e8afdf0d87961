#!/bin/bash
# GPU call A of round 2: the full -m gpu suite (incl. the BASELINE-shape goldens) and the default bench line (decode + SCST).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_gputest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r02a_gputest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
