#!/bin/bash
# GPU call N of round 2 (gpurun --gpus 2): two-rank gradient-sync check with the step replayed as a graph (external event nodes), DataParallel tests,
# 2-rank and 1-rank bench lines.
set -u
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/r02n_$name.log 2>&1; echo "== $name rc=$? :: $(tail -1 gpurun_out/r02n_$name.log | cut -c1-200)"; grep -E "^(FAILED|ERROR)" gpurun_out/r02n_$name.log | head -20; }
run multi 1200 python -m pytest tests/test_gpu_multi.py -q -m gpu
grep -E "^E  |SYNC-OK|differ" gpurun_out/r02n_multi.log | head -12 | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
CAPB200_GRAPH_DEBUG=1 timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02n_bench_2gpu.json 2> gpurun_out/r02n_bench_2gpu.err; echo "bench2 rc=$?"
grep -i "capb200:" gpurun_out/r02n_bench_2gpu.err | head -4
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02n_bench_1gpu.json 2> gpurun_out/r02n_bench_1gpu.err; echo "bench1 rc=$?"
python - <<'PY'
import json
for n in ('1gpu', '2gpu'):
    try:
        d = json.loads(open('gpurun_out/r02n_bench_%s.json' % n).read().strip().splitlines()[-1])
    except Exception as e:
        print(n, 'unreadable', e); continue
    s = d.get('scst') or {}
    print(n, 'decode', round(d['value']), 'ms', d.get('per_rank_ms_per_step'), '| scst', round(s.get('value', 0)), 'ms', s.get('per_rank_ms_per_step'), 'allreduce', s.get('allreduce_ms'))
PY
CAPB200_SCST_GRAPH_SYNC=0 timeout 600 $TR bench.py --gpus 2 --workload aoa_scst --steps 20 --warmup 5 > gpurun_out/r02n_bench_2gpu_scst_eager_sync.json 2> gpurun_out/r02n_bench_2gpu_scst_eager_sync.err; echo "scst2 eager-with-listener rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r02n_bench_2gpu_scst_eager_sync.json').read().strip().splitlines()[-1]); print('2gpu scst, eager steps when a listener is registered:', round(d['value']), d['per_rank_ms_per_step'], d['allreduce_ms'])"
