#!/bin/bash
# GPU call A of round 2: every test file in its own process (a trap in one kernel must not take the other files down), the new tcgen05 tf32
# GEMM first; then the default bench line (decode + SCST) and A/B runs of the switchable pieces.
set -u
mkdir -p gpurun_out
run() { # name, timeout, command...
  local name=$1 to=$2; shift 2
  timeout $to "$@" > gpurun_out/r02a_$name.log 2>&1
  echo "== $name rc=$? :: $(tail -1 gpurun_out/r02a_$name.log | cut -c1-200)"
}
run tf32 600 python -m pytest tests/test_gpu_ops.py -q -x -k "tf32x3"
for f in tests/test_gpu_ops.py tests/test_gpu_decode.py tests/test_gpu_options.py tests/test_gpu_aoa.py tests/test_gpu_transformer.py tests/test_gpu_scst.py tests/test_gpu_baseline_shapes.py tests/test_gpu_multi.py; do
  n=$(basename $f .py)
  run $n 1500 python -m pytest $f -q -m gpu
  grep -E "^(FAILED|ERROR)" gpurun_out/r02a_$n.log | head -20
done
# if the tcgen05 training GEMMs misbehave, the same training tests on the legacy mma.sync path separate kernel bugs from the rest
if ! tail -1 gpurun_out/r02a_test_gpu_scst.log | grep -q " passed" || tail -1 gpurun_out/r02a_test_gpu_scst.log | grep -q "failed"; then
  CAPB200_SKINNY_LEGACY=1 run scst_legacy 1500 python -m pytest tests/test_gpu_scst.py -q -m gpu
  grep -E "^(FAILED|ERROR)" gpurun_out/r02a_scst_legacy.log | head -20
fi
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
CAPB200_SKINNY_LEGACY=1 CAPB200_SCST_SERIAL_GREEDY=1 timeout 600 python bench.py --workload aoa_scst --steps 20 --warmup 5 > gpurun_out/r02a_bench_scst_legacy.json 2> gpurun_out/r02a_bench_scst_legacy.err
echo "legacy scst rc=$?"; tail -c 900 gpurun_out/r02a_bench_scst_legacy.json
CAPB200_SPLIT_LANG=1 CAPB200_BENCH_NO_SCST=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench_split.json 2> gpurun_out/r02a_bench_split.err
echo "split-lang rc=$?"; tail -c 1500 gpurun_out/r02a_bench_split.json | cut -c1-700
# warm per-kernel tables (CUPTI) of the AoANet SCST step and the headline decode
timeout 600 python tools/scst_table.py 10 aoa > gpurun_out/r02a_scst_table_aoa.txt 2>&1; echo "scst table rc=$?"; head -30 gpurun_out/r02a_scst_table_aoa.txt
timeout 600 python tools/kernel_table.py 256 5 updown > gpurun_out/r02a_kernel_table.txt 2>&1; echo "decode table rc=$?"; head -16 gpurun_out/r02a_kernel_table.txt
