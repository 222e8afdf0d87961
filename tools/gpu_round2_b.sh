#!/bin/bash
# GPU call B of round 2: the files that failed in call A (test bugs fixed), the bench line, the SCST A/B and the kernel tables.
set -u
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/r02b_$name.log 2>&1; echo "== $name rc=$? :: $(tail -1 gpurun_out/r02b_$name.log | cut -c1-200)"; grep -E "^(FAILED|ERROR)" gpurun_out/r02b_$name.log | head -20; }
run tf32 600 python -m pytest tests/test_gpu_ops.py -q -k "tf32x3"
run scst 1500 python -m pytest tests/test_gpu_scst.py -q -m gpu
run misc 900 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_decode.py -q -m gpu -k "aoa_scst_step_at_config or scst_forward_values"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
echo "bench rc=$?"; tail -c 3500 gpurun_out/r02b_bench.json; tail -5 gpurun_out/r02b_bench.err
CAPB200_SKINNY_LEGACY=1 CAPB200_SCST_SERIAL_GREEDY=1 timeout 600 python bench.py --workload aoa_scst --steps 20 --warmup 5 > gpurun_out/r02b_bench_scst_legacy.json 2> gpurun_out/r02b_bench_scst_legacy.err
echo "legacy scst rc=$?"; tail -c 700 gpurun_out/r02b_bench_scst_legacy.json | cut -c1-700
CAPB200_SCST_SERIAL_GREEDY=1 timeout 600 python bench.py --workload aoa_scst --steps 20 --warmup 5 > gpurun_out/r02b_bench_scst_serial.json 2> gpurun_out/r02b_bench_scst_serial.err
echo "serial-greedy scst rc=$?"; tail -c 700 gpurun_out/r02b_bench_scst_serial.json | cut -c1-400
timeout 600 python bench.py --workload updown_scst --steps 20 --warmup 5 > gpurun_out/r02b_bench_updown_scst.json 2> gpurun_out/r02b_bench_updown_scst.err
echo "updown scst rc=$?"; tail -c 700 gpurun_out/r02b_bench_updown_scst.json | cut -c1-400
timeout 600 python tools/scst_table.py 10 aoa > gpurun_out/r02b_scst_table_aoa.txt 2>&1; echo "scst table rc=$?"; head -34 gpurun_out/r02b_scst_table_aoa.txt | cut -c1-200
