#!/bin/bash
# per-member hand-off between the quad spec->grid kernel and the column kernel: parity and timing with / without
tag=${1:-r2x}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_kernel_variants_gpu.py tests/test_ensemble_gpu.py tests/test_output_gpu.py -x -q -k "48h or ensemble or partition or small_blocks or restart or sppt" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
for v in 1 0; do for m in 8 16; do echo -n "member_ready=$v m$m: "; SPEEDY_MEMBER_READY=$v timeout 200 python tools/ktime.py $m 2>&1 | tail -1 | cut -c1-100; done; done
for v in 1 0; do echo -n "member_ready=$v sppt m8: "; SPEEDY_MEMBER_READY=$v timeout 200 python tools/ktime.py 8 30 sppt 2>&1 | tail -1 | cut -c1-100; done
