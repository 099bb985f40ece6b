#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernel_variants_gpu.py -x -q -k "member_ready" 2>&1 | tail -3
for cfg in "8 1" "4 2" "2 4" "16 1" "8 2" "4 4" "8 1 1" "4 2 1"; do timeout 200 python tools/two_ctx.py $cfg 2>&1 | tail -1; done
