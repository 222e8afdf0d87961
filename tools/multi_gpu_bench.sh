#!/bin/bash
# GPU box (gpurun --gpus N): the driver's torchrun launch of bench.py for the headline and the SCST workloads
N=${1:-2}
mkdir -p gpurun_out
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" 2>&1 | grep '"metric"' | tail -1 > gpurun_out/$2
  cut -c1-330 gpurun_out/$2
}
run 29511 bench_${N}gpu_updown_beam.json --steps 20 --warmup 5
run 29512 bench_${N}gpu_aoa_scst.json --workload aoa_scst --steps 10 --warmup 3
run 29513 bench_${N}gpu_updown_scst.json --workload updown_scst --steps 10 --warmup 3
run 29514 bench_${N}gpu_tfm_beam.json --workload transformer_beam --batch 64 --steps 10 --warmup 3
