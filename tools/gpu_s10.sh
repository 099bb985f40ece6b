#!/bin/bash
tag=${1:-r2x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernel_variants_gpu.py tests/test_ensemble_gpu.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
for m in 8 16; do SPEEDY_TRACE_STAMPS=1 timeout 200 python tools/ktime.py $m 2>&1 | tail -8 | cut -c1-330; done
timeout 200 python tools/ktime.py 8 30 sppt 2>&1 | tail -1 | cut -c1-100
timeout 300 python bench_transforms.py 2>&1 | tail -4 | cut -c1-400
