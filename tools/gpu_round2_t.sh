#!/bin/bash
# GPU call R of round 2 (gpurun --gpus 8): the default bench line on eight ranks (decode: no collective; SCST: overlapped all-reduce inside the step graph).
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02t_bench_8gpu.json 2> gpurun_out/r02t_bench_8gpu.err; echo "bench4 rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02t_bench_8gpu.json').read().strip().splitlines()[-1])
s = d.get('scst') or {}
print('8gpu decode', round(d['value']), 'per-rank ms', [round(v, 3) for v in d.get('per_rank_ms_per_step', [])], d['clocks'])
print('8gpu scst', round(s.get('value', 0)), 'per-rank ms', [round(v, 3) for v in s.get('per_rank_ms_per_step', [])], 'allreduce exposed', s.get('allreduce_ms'), s.get('clocks'))
PY
tail -3 gpurun_out/r02t_bench_8gpu.err | cut -c1-200
