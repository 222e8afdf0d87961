#!/bin/bash
# GPU call A of round 2: the new tcgen05 tf32 GEMM tests first, then the full -m gpu suite (incl. the BASELINE-shape goldens), then the
# default bench line (decode + SCST) with the tcgen05 training GEMMs and, for comparison, with the legacy mma.sync ones.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -k "tf32x3" > gpurun_out/r02a_tf32.log 2>&1
echo "tf32 ops rc=$?"; tail -12 gpurun_out/r02a_tf32.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02a_gputest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r02a_gputest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
CAPB200_SKINNY_LEGACY=1 timeout 600 python bench.py --workload aoa_scst --steps 20 --warmup 5 > gpurun_out/r02a_bench_scst_legacy.json 2> gpurun_out/r02a_bench_scst_legacy.err
echo "legacy scst rc=$?"; tail -c 1200 gpurun_out/r02a_bench_scst_legacy.json
