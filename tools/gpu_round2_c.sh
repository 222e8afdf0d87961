#!/bin/bash
# GPU call C of round 2: tests of the rewritten training-step kernels, tf32 sweep / truncation variant, GEMM phase trace, bench lines.
set -u
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/r02c_$name.log 2>&1; echo "== $name rc=$? :: $(tail -1 gpurun_out/r02c_$name.log | cut -c1-200)"; grep -E "^(FAILED|ERROR)" gpurun_out/r02c_$name.log | head -20; }
run scst 1500 python -m pytest tests/test_gpu_scst.py tests/test_gpu_aoa.py tests/test_gpu_transformer.py -q -m gpu
run decode 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_baseline_shapes.py tests/test_gpu_options.py -q -m gpu
CAPB200_TF32_TRUNC=1 run tf32_trunc 600 python -m pytest tests/test_gpu_ops.py -q -k "tf32x3"
timeout 300 python tools/tf32_sweep.py 50 > gpurun_out/r02c_tf32_sweep.txt 2>&1; echo "sweep rc=$?"; cat gpurun_out/r02c_tf32_sweep.txt
CAPB200_TF32_TRUNC=1 timeout 300 python tools/tf32_sweep.py 50 > gpurun_out/r02c_tf32_sweep_trunc.txt 2>&1; echo "sweep trunc rc=$?"; cat gpurun_out/r02c_tf32_sweep_trunc.txt
for s in "1280 4000 3000" "1280 4000 2000" "1280 9488 1000"; do timeout 200 python tools/gemm_trace.py $s >> gpurun_out/r02c_gemm_trace.txt 2>&1; done; cat gpurun_out/r02c_gemm_trace.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c_bench.json').read().strip().splitlines()[-1])
print('decode', round(d['value']), 'cap/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), '; gemm frac', round(d['roofline']['frac'],3), d['roofline']['per_gemm_ms_per_step'])
s=d['scst']; print('scst', round(s['value']), 'samples/s', round(s['ms_per_step'],2), 'ms', s['step_wall_ms'], 'launches', s['launches'])
PY
tail -3 gpurun_out/r02c_bench.err
CAPB200_BENCH_STEP_TIMES=1 timeout 600 python bench.py --workload aoa_scst --steps 20 --warmup 5 > gpurun_out/r02c_bench_aoa_scst.json 2> gpurun_out/r02c_bench_aoa_scst.err
echo "standalone scst rc=$?"; grep "per-step wall" gpurun_out/r02c_bench_aoa_scst.err; tail -c 600 gpurun_out/r02c_bench_aoa_scst.json | cut -c1-300
timeout 600 python tools/scst_table.py 10 aoa > gpurun_out/r02c_scst_table_aoa.txt 2>&1; echo "scst table rc=$?"; head -30 gpurun_out/r02c_scst_table_aoa.txt | cut -c1-190
