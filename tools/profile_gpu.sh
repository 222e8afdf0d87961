#!/bin/bash
# Runs on the GPU box (under gpurun): launch list + one full ncu capture of the GEMM kernels of one timestep.
# Outputs land in gpurun_out/; summaries are copied to profiles/ by tools/summarise_profiles.py in the build container.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
# (1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "launch list rc=$?"
# (2) full capture of the four GEMMs of one timestep at N = B*beam rows (att_lstm, h2att, lang_lstm, logit)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 12 -c 5 -o gpurun_out/${TAG}_gemm \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
# (3) full capture of the step's non-GEMM kernels (one launch each, steady state)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'vocab_stats|att_score|att_combine|state_gather|beam_step' -s 25 -c 5 -o gpurun_out/${TAG}_small \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_small.log 2>&1
echo "small capture rc=$?"
