#!/bin/bash
# GPU call H of round 2: SCST step as a CUDA graph (seed salt), no host sync on re-binds, greedy enqueue after the prologue; fp16 range guard test;
# compute-sanitizer passes.
set -u
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/r02h_$name.log 2>&1; echo "== $name rc=$? :: $(tail -1 gpurun_out/r02h_$name.log | cut -c1-200)"; grep -E "^(FAILED|ERROR)" gpurun_out/r02h_$name.log | head -20; }
run scst 1500 python -m pytest tests/test_gpu_scst.py tests/test_gpu_aoa.py tests/test_gpu_tfm_train.py -q -m gpu
grep -E "^E  " gpurun_out/r02h_scst.log | head -12 | cut -c1-300
run decode 1200 python -m pytest tests/test_gpu_decode.py tests/test_gpu_baseline_shapes.py tests/test_gpu_transformer.py tests/test_gpu_options.py -q -m gpu
grep -E "^E  " gpurun_out/r02h_decode.log | head -12 | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02h_bench.json').read().strip().splitlines()[-1])
print('decode', round(d['value']), 'cap/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), '; gemm frac', round(d['roofline']['frac'],3))
s=d['scst']; print('scst', round(s['value']), 'samples/s', round(s['ms_per_step'],2), 'ms', s['step_wall_ms'], 'launches', s['launches'])
PY
tail -3 gpurun_out/r02h_bench.err
CAPB200_SCST_GRAPH=0 timeout 600 python bench.py --workload aoa_scst --steps 20 --warmup 5 > gpurun_out/r02h_bench_aoa_scst_eager.json 2> gpurun_out/r02h_bench_aoa_scst_eager.err; echo "eager scst rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r02h_bench_aoa_scst_eager.json').read().strip().splitlines()[-1]); print('aoa scst eager (no graph)', round(d['value']), round(d['ms_per_step'],2), d['step_wall_ms'])"
timeout 600 python bench.py --workload updown_scst --steps 20 --warmup 5 > gpurun_out/r02h_bench_updown_scst.json 2> gpurun_out/r02h_bench_updown_scst.err; python -c "
import json; d=json.loads(open('gpurun_out/r02h_bench_updown_scst.json').read().strip().splitlines()[-1]); print('updown scst', round(d['value']), round(d['ms_per_step'],2))"
timeout 600 python bench.py --workload transformer_scst --steps 10 --warmup 3 > gpurun_out/r02h_bench_tfm_scst.json 2> gpurun_out/r02h_bench_tfm_scst.err; python -c "
import json; d=json.loads(open('gpurun_out/r02h_bench_tfm_scst.json').read().strip().splitlines()[-1]); print('transformer scst', round(d['value']), round(d['ms_per_step'],2), d['launches'])"
timeout 600 python tools/scst_timeline.py aoa gpurun_out/r02h_timeline_aoa.json > gpurun_out/r02h_timeline_aoa.txt 2>&1; echo "timeline rc=$?"; grep -v Warn gpurun_out/r02h_timeline_aoa.txt | cut -c1-200 | head -30
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -x -k "tf32x3_tcgen05_linear" > gpurun_out/r02h_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r02h_racecheck.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_scst.py -q -x -k "test_aoa_scst_step_gradients" > gpurun_out/r02h_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r02h_memcheck.log | cut -c1-200
