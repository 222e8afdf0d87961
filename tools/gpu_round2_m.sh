#!/bin/bash
# GPU call M (attention tiling, 1024-thread sampling, LSTM+LN) of round 2: fused AoANet loop kernels, trace-free decode GEMM, graph debug counters.
set -u
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/r02m_$name.log 2>&1; echo "== $name rc=$? :: $(tail -1 gpurun_out/r02m_$name.log | cut -c1-200)"; grep -E "^(FAILED|ERROR)" gpurun_out/r02m_$name.log | head -20; }
run scst 1500 python -m pytest tests/test_gpu_scst.py tests/test_gpu_aoa.py tests/test_gpu_tfm_train.py -q -m gpu
grep -E "^E  " gpurun_out/r02m_scst.log | head -12 | cut -c1-300
run shapes 1500 python -m pytest tests/test_gpu_baseline_shapes.py -q -m gpu -s
grep -E "worst relative|^E  " gpurun_out/r02m_shapes.log | head | cut -c1-300
CAPB200_GRAPH_DEBUG=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02m_bench.json').read().strip().splitlines()[-1])
print('decode', round(d['value']), 'cap/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), '; gemm frac', round(d['roofline']['frac'],3), d['roofline']['per_gemm_ms_per_step'])
s=d['scst']; print('scst', round(s['value']), 'samples/s', round(s['ms_per_step'],2), 'ms', s['step_wall_ms'], 'launches', s['launches'])
PY
grep -i "capb200:" gpurun_out/r02m_bench.err | head; tail -2 gpurun_out/r02m_bench.err
timeout 600 python tools/scst_timeline.py aoa gpurun_out/r02m_timeline_aoa.json > gpurun_out/r02m_timeline_aoa.txt 2>&1; echo "timeline rc=$?"; grep -v Warn gpurun_out/r02m_timeline_aoa.txt | cut -c1-200 | head -12
CAPB200_GRAPH_DEBUG=1 timeout 600 python bench.py --workload updown_scst --steps 20 --warmup 5 > gpurun_out/r02m_bench_updown_scst.json 2> gpurun_out/r02m_bench_updown_scst.err; python -c "
import json; d=json.loads(open('gpurun_out/r02m_bench_updown_scst.json').read().strip().splitlines()[-1]); print('updown scst', round(d['value']), round(d['ms_per_step'],2))"; grep -i "capb200:" gpurun_out/r02m_bench_updown_scst.err | head -3
CAPB200_GRAPH_DEBUG=1 timeout 600 python bench.py --workload transformer_scst --steps 10 --warmup 4 > gpurun_out/r02m_bench_tfm_scst.json 2> gpurun_out/r02m_bench_tfm_scst.err; python -c "
import json; d=json.loads(open('gpurun_out/r02m_bench_tfm_scst.json').read().strip().splitlines()[-1]); print('transformer scst', round(d['value']), round(d['ms_per_step'],2), d['launches'])"; grep -i "capb200:" gpurun_out/r02m_bench_tfm_scst.err | head -3
