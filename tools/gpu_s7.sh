#!/bin/bash
# batch column kernels: parity (48 h with 8 members vs oracle, ensembles, restart) and timing with / without the split
tag=${1:-r2x}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernel_variants_gpu.py tests/test_ensemble_gpu.py tests/test_output_gpu.py -x -q -k "48h or ensemble or partition or small_blocks or restart or sppt" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
for v in 1 0; do for m in 8 16; do echo -n "col_split=$v m$m: "; SPEEDY_COL_SPLIT=$v timeout 200 python tools/ktime.py $m 2>&1 | tail -2 | cut -c1-215 | tr '\n' ' '; echo; done; done
