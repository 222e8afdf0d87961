#!/bin/bash
# GPU call P of round 2: split-K through a global scratch buffer (last-arriver reduction) vs the cluster form, fused additive attention of the decode step.
set -u
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/r02p_$name.log 2>&1; echo "== $name rc=$? :: $(tail -1 gpurun_out/r02p_$name.log | cut -c1-200)"; grep -E "^(FAILED|ERROR)" gpurun_out/r02p_$name.log | head -20; }
run ops 900 python -m pytest tests/test_gpu_ops.py -q
grep -E "^E  " gpurun_out/r02p_ops.log | head -8 | cut -c1-300
timeout 300 python tools/tf32_sweep.py 50 > gpurun_out/r02p_tf32_sweep.txt 2>&1; echo "sweep rc=$?"; cat gpurun_out/r02p_tf32_sweep.txt
CAPB200_TF32_CLUSTER=1 timeout 300 python tools/tf32_sweep.py 50 > gpurun_out/r02p_tf32_sweep_cluster.txt 2>&1; echo "sweep cluster rc=$?"; cat gpurun_out/r02p_tf32_sweep_cluster.txt
run scst 1500 python -m pytest tests/test_gpu_scst.py tests/test_gpu_aoa.py tests/test_gpu_tfm_train.py -q -m gpu
grep -E "^E  " gpurun_out/r02p_scst.log | head -8 | cut -c1-300
run decode 1500 python -m pytest tests/test_gpu_decode.py tests/test_gpu_baseline_shapes.py tests/test_gpu_options.py tests/test_gpu_transformer.py -q -m gpu
grep -E "^E  " gpurun_out/r02p_decode.log | head -8 | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02p_bench.json').read().strip().splitlines()[-1])
print('decode', round(d['value']), 'cap/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), '; gemm frac', round(d['roofline']['frac'],3))
s=d['scst']; print('scst', round(s['value']), 'samples/s', round(s['ms_per_step'],2), 'ms', s['step_wall_ms'], 'launches', s['launches'])
PY
tail -2 gpurun_out/r02p_bench.err
CAPB200_TF32_CLUSTER=1 CAPB200_ATT_SPLIT=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02p_bench_ab.json 2> gpurun_out/r02p_bench_ab.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02p_bench_ab.json').read().strip().splitlines()[-1])
print('A/B (cluster split-K, split attention kernels): decode', round(d['value']), round(d['ms_per_step'],3), 'ms | scst', round(d['scst']['value']), round(d['scst']['ms_per_step'],2), 'ms')
PY
