#!/bin/bash
# GPU call F of round 2: fused Adam, pinned reference upload, chunked self-attention kernels, eight converter warps in the tf32 GEMM.
set -u
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/r02f_$name.log 2>&1; echo "== $name rc=$? :: $(tail -1 gpurun_out/r02f_$name.log | cut -c1-200)"; grep -E "^(FAILED|ERROR)" gpurun_out/r02f_$name.log | head -20; }
run ops 900 python -m pytest tests/test_gpu_ops.py -q -k "tf32x3 or adam"
run scst 1500 python -m pytest tests/test_gpu_scst.py tests/test_gpu_aoa.py tests/test_gpu_transformer.py tests/test_gpu_baseline_shapes.py -q -m gpu
timeout 300 python tools/tf32_sweep.py 50 > gpurun_out/r02f_tf32_sweep.txt 2>&1; echo "sweep rc=$?"; cat gpurun_out/r02f_tf32_sweep.txt
timeout 600 python tools/scst_timeline.py aoa gpurun_out/r02f_timeline_aoa.json > gpurun_out/r02f_timeline_aoa.txt 2>&1; echo "timeline rc=$?"; grep -v Warn gpurun_out/r02f_timeline_aoa.txt | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench.json').read().strip().splitlines()[-1])
print('decode', round(d['value']), 'cap/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), '; gemm frac', round(d['roofline']['frac'],3))
s=d['scst']; print('scst', round(s['value']), 'samples/s', round(s['ms_per_step'],2), 'ms', s['step_wall_ms'], 'launches', s['launches'])
PY
tail -3 gpurun_out/r02f_bench.err
timeout 600 python bench.py --workload updown_scst --steps 20 --warmup 5 > gpurun_out/r02f_bench_updown_scst.json 2> gpurun_out/r02f_bench_updown_scst.err; echo "updown scst rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02f_bench_updown_scst.json').read().strip().splitlines()[-1]); print('updown scst', round(d['value']), round(d['ms_per_step'],2))"
