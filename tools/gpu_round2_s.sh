#!/bin/bash
# GPU call S of round 2: hybrid split-K policy of the tf32 GEMM (cluster form, except 16 clusters of 8 -> independent K-ranks through scratch).
set -u
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/r02s_$name.log 2>&1; echo "== $name rc=$? :: $(tail -1 gpurun_out/r02s_$name.log | cut -c1-200)"; grep -E "^(FAILED|ERROR)" gpurun_out/r02s_$name.log | head -20; }
run ops 900 python -m pytest tests/test_gpu_ops.py -q
run scst 1500 python -m pytest tests/test_gpu_scst.py tests/test_gpu_aoa.py tests/test_gpu_tfm_train.py tests/test_gpu_baseline_shapes.py -q -m gpu
grep -E "^E  " gpurun_out/r02s_scst.log | head -8 | cut -c1-300
timeout 300 python tools/tf32_sweep.py 50 > gpurun_out/r02s_tf32_sweep.txt 2>&1; echo "sweep rc=$?"; cat gpurun_out/r02s_tf32_sweep.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02s_bench.json').read().strip().splitlines()[-1])
print('decode', round(d['value']), 'cap/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']))
s=d['scst']; print('scst', round(s['value']), 'samples/s', round(s['ms_per_step'],2), 'ms', s['step_wall_ms'], 'launches', s['launches'])
PY
