#!/bin/bash
# GPU call I of round 2: PDL probe, graph replay tests, config-size Transformer training parity, sync check needs 2 GPUs (skipped here).
set -u
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/r02i_$name.log 2>&1; echo "== $name rc=$? :: $(tail -1 gpurun_out/r02i_$name.log | cut -c1-200)"; grep -E "^(FAILED|ERROR)" gpurun_out/r02i_$name.log | head -20; }
timeout 120 tools/probes/pdl_probe > gpurun_out/r02i_pdl_probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r02i_pdl_probe.txt
run graph 900 python -m pytest tests/test_gpu_scst.py -q -m gpu -k "graph_replay or loss_wrapper"
grep -E "^E  " gpurun_out/r02i_graph.log | head -12 | cut -c1-300
run shapes 900 python -m pytest tests/test_gpu_baseline_shapes.py -q -m gpu -s -k "transformer_training"
grep -E "worst relative|^E  " gpurun_out/r02i_shapes.log | head | cut -c1-300
run range 600 python -m pytest tests/test_gpu_decode.py -q -m gpu -k "range_guard"
grep -E "^E  " gpurun_out/r02i_range.log | head -12 | cut -c1-300
