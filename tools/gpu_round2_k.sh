#!/bin/bash
# GPU call K of round 2: ncu evidence -- launch lists of one decode batch and one AoANet SCST step, full captures of the two dominant kernels.
set -u
mkdir -p gpurun_out
export CAPB200_BENCH_NO_SCST=1
# (1) launch list of the decode bench command (one warm batch skipped: -s = launches of bind + first batch is unknown, so list everything of 2 steps)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02k_launches_decode.csv \
    python bench.py --steps 1 --warmup 3 > gpurun_out/r02k_ncu_decode.log 2>&1; echo "ncu decode list rc=$?"
# (2) full capture of the dominant decode kernel (lang_lstm / att_lstm / logit pair GEMMs): skip the first 30 instances (warm-up batches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair_kernel -s 186 -c 3 -o gpurun_out/r02k_gemm_pair \
    python bench.py --steps 1 --warmup 3 > gpurun_out/r02k_ncu_pair.log 2>&1; echo "ncu pair rc=$?"
unset CAPB200_BENCH_NO_SCST
# (3) launch list of one eager AoANet SCST step (graphs off so that every kernel is a separate launch for ncu either way)
CAPB200_SCST_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5000 -c 1700 --csv --log-file gpurun_out/r02k_launches_scst.csv \
    python bench.py --workload aoa_scst --steps 1 --warmup 3 > gpurun_out/r02k_ncu_scst.log 2>&1; echo "ncu scst list rc=$?"
# (4) full capture of the skinny tf32 GEMM (sampling-loop call sites)
CAPB200_SCST_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3_kernel -s 600 -c 4 -o gpurun_out/r02k_gemm_tf32 \
    python bench.py --workload aoa_scst --steps 1 --warmup 3 > gpurun_out/r02k_ncu_tf32.log 2>&1; echo "ncu tf32 rc=$?"
ls -la gpurun_out/r02k_* | head; tail -2 gpurun_out/r02k_ncu_pair.log | cut -c1-200
