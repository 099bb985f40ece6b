#!/bin/bash
tag=${1:-r2x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernel_variants_gpu.py tests/test_ensemble_gpu.py tests/test_output_gpu.py tests/test_physics_states_gpu.py -x -q -k "48h or ensemble or partition or small_blocks or restart or sppt or plumbing or states" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
for m in 6 8 12 16; do SPEEDY_DEBUG_OCC=1 SPEEDY_TRACE_STAMPS=1 timeout 200 python tools/ktime.py $m 2>&1 | grep "CTAs per SM\|column stamps\|in-graph\|members" | cut -c1-300; done
timeout 200 python tools/ktime.py 8 30 sppt 2>&1 | tail -1 | cut -c1-100
