#!/bin/bash
tag=${1:-r2x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernel_variants_gpu.py tests/test_ensemble_gpu.py tests/test_output_gpu.py tests/test_member_counts_gpu.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
for v in 1 0; do echo -n "sppt_fold=$v sppt m8: "; SPEEDY_SPPT_FOLD=$v timeout 200 python tools/ktime.py 8 30 sppt 2>&1 | tail -1 | cut -c1-100; echo -n "sppt_fold=$v sppt m16: "; SPEEDY_SPPT_FOLD=$v timeout 200 python tools/ktime.py 16 30 sppt 2>&1 | tail -1 | cut -c1-100; done
timeout 200 python tools/ktime.py 8 2>&1 | tail -1 | cut -c1-100
