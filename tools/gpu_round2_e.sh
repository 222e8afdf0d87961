#!/bin/bash
# GPU call E of round 2: truncating tf32 split as the default (accuracy + config-size parity), SCST step timeline.
set -u
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout $to "$@" > gpurun_out/r02e_$name.log 2>&1; echo "== $name rc=$? :: $(tail -1 gpurun_out/r02e_$name.log | cut -c1-200)"; grep -E "^(FAILED|ERROR)" gpurun_out/r02e_$name.log | head -20; }
run tf32 600 python -m pytest tests/test_gpu_ops.py -q -k "tf32x3"
run scst 1500 python -m pytest tests/test_gpu_scst.py tests/test_gpu_baseline_shapes.py -q -m gpu
timeout 600 python tools/scst_timeline.py aoa gpurun_out/r02e_timeline_aoa.json > gpurun_out/r02e_timeline_aoa.txt 2>&1; echo "timeline rc=$?"; cat gpurun_out/r02e_timeline_aoa.txt | cut -c1-220
timeout 600 python tools/scst_timeline.py updown > gpurun_out/r02e_timeline_updown.txt 2>&1; echo "timeline rc=$?"; cat gpurun_out/r02e_timeline_updown.txt | cut -c1-220
