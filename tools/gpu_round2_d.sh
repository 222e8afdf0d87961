#!/bin/bash
# GPU call D of round 2 (gpurun --gpus 2): DataParallel tests, 2-rank bench lines (decode + SCST with overlapped gradient all-reduce),
# and one compute-sanitizer racecheck pass over the tcgen05 kernels on GPU 0.
set -u
mkdir -p gpurun_out
nvidia-smi -L
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/r02d_$name.log 2>&1; echo "== $name rc=$? :: $(tail -1 gpurun_out/r02d_$name.log | cut -c1-200)"; grep -E "^(FAILED|ERROR)" gpurun_out/r02d_$name.log | head -20; }
run multi 900 python -m pytest tests/test_gpu_multi.py -q -m gpu
run scst1 900 python -m pytest tests/test_gpu_scst.py tests/test_gpu_baseline_shapes.py tests/test_gpu_tfm_train.py -q -m gpu -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02d_bench_2gpu.json 2> gpurun_out/r02d_bench_2gpu.err; echo "bench2 rc=$?"
tail -c 1800 gpurun_out/r02d_bench_2gpu.json; echo; tail -3 gpurun_out/r02d_bench_2gpu.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02d_bench_1gpu.json 2> gpurun_out/r02d_bench_1gpu.err; echo "bench1 rc=$?"
python - <<'PY'
import json
for n in ('1gpu', '2gpu'):
    try:
        d = json.loads(open('gpurun_out/r02d_bench_%s.json' % n).read().strip().splitlines()[-1])
    except Exception as e:
        print(n, 'unreadable', e); continue
    s = d.get('scst') or {}
    print(n, 'decode', round(d['value']), 'ms', d.get('per_rank_ms_per_step'), '| scst', round(s.get('value', 0)), 'ms', s.get('per_rank_ms_per_step'), 'allreduce', s.get('allreduce_ms'))
PY
CAPB200_SCST_NO_OVERLAP=1 timeout 600 $TR bench.py --gpus 2 --workload aoa_scst --steps 20 --warmup 5 > gpurun_out/r02d_bench_2gpu_scst_nooverlap.json 2> gpurun_out/r02d_bench_2gpu_scst_nooverlap.err; echo "scst2 no-overlap rc=$?"
tail -c 700 gpurun_out/r02d_bench_2gpu_scst_nooverlap.json; echo
timeout 600 python bench.py --workload transformer_scst --steps 10 --warmup 3 > gpurun_out/r02d_bench_tfm_scst.json 2> gpurun_out/r02d_bench_tfm_scst.err; echo "tfm scst rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02d_bench_tfm_scst.json').read().strip().splitlines()[-1]); print('transformer scst', round(d['value']), 'samples/s', round(d['ms_per_step'],2), 'ms', d['launches'], 'launches')"
# racecheck / memcheck of the tensor-core kernels (one GPU, small cases)
export CUDA_VISIBLE_DEVICES=0
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -x -k "tf32x3_tcgen05_forward or tc_linear_small" > gpurun_out/r02d_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -8 gpurun_out/r02d_racecheck.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_scst.py -q -x -k "aoa_scst_step_matches" > gpurun_out/r02d_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -8 gpurun_out/r02d_memcheck.log | cut -c1-200
