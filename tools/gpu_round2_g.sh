#!/bin/bash
# GPU call G of round 2: Transformer training steps (XE + SCST), fused Adam arithmetic fix, direct gradient views.
set -u
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/r02g_$name.log 2>&1; echo "== $name rc=$? :: $(tail -1 gpurun_out/r02g_$name.log | cut -c1-200)"; grep -E "^(FAILED|ERROR)" gpurun_out/r02g_$name.log | head -20; }
run tfm_train 1200 python -m pytest tests/test_gpu_tfm_train.py -q -m gpu -x
grep -E "^E  " gpurun_out/r02g_tfm_train.log | head -12 | cut -c1-300
run ops 900 python -m pytest tests/test_gpu_ops.py -q -k "adam"
run scst 1500 python -m pytest tests/test_gpu_scst.py tests/test_gpu_aoa.py tests/test_gpu_transformer.py -q -m gpu
run shapes 1500 python -m pytest tests/test_gpu_baseline_shapes.py -q -m gpu -s
grep -E "worst relative|^E  " gpurun_out/r02g_shapes.log | head | cut -c1-300
timeout 600 python tools/scst_timeline.py aoa gpurun_out/r02g_timeline_aoa.json > gpurun_out/r02g_timeline_aoa.txt 2>&1; echo "timeline rc=$?"; grep -v Warn gpurun_out/r02g_timeline_aoa.txt | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02g_bench.json').read().strip().splitlines()[-1])
print('decode', round(d['value']), 'cap/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), '; gemm frac', round(d['roofline']['frac'],3))
s=d['scst']; print('scst', round(s['value']), 'samples/s', round(s['ms_per_step'],2), 'ms', s['step_wall_ms'], 'launches', s['launches'])
PY
tail -3 gpurun_out/r02g_bench.err
