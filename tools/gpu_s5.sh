#!/bin/bash
tag=${1:-r2x}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernel_variants_gpu.py -x -q -k "48h" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python -m pytest tests/test_model_gpu.py tests/test_ensemble_gpu.py -x -q > gpurun_out/${tag}_pytest2.log 2>&1; echo "pytest2 rc=$?"; tail -4 gpurun_out/${tag}_pytest2.log
for m in 8 16; do timeout 200 python tools/ktime.py $m 2>&1 | tail -2; done
timeout 200 python tools/ktime.py 1 47 2>&1 | tail -1
