#!/bin/bash
# GPU call Q of round 2: what the driver runs at round end -- the whole GPU suite, smoke(), both bench arms.
set -u
mkdir -p gpurun_out
timeout 2000 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02q_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? :: $(tail -1 gpurun_out/r02q_pytest_gpu.log | cut -c1-200)"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r02q_pytest_gpu.log | head -10 | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02q_smoke.log 2>&1; echo "smoke rc=$? :: $(tail -1 gpurun_out/r02q_smoke.log | cut -c1-200)"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02q_bench_reference.json 2> gpurun_out/r02q_bench_reference.err; echo "reference arm rc=$?"; tail -c 600 gpurun_out/r02q_bench_reference.json; echo
timeout 900 python bench.py > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02q_bench.json').read().strip().splitlines()[-1])
print('decode', round(d['value']), 'cap/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), '; gemm frac', round(d['roofline']['frac'],3), 'clocks', d['clocks'])
s=d['scst']; print('scst', round(s['value']), 'samples/s', round(s['ms_per_step'],2), 'ms', s['step_wall_ms'], 'launches', s['launches'], s['clocks'])
PY
tail -2 gpurun_out/r02q_bench.err
