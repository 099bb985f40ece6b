#!/bin/bash
# round-2 session 1: validate the whole-field grid->spec kernel and time it against the default
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernel_variants_gpu.py -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log
timeout 300 python tools/bench_transforms.py 30 _base > gpurun_out/r2a_xf_base.log 2>&1; tail -10 gpurun_out/r2a_xf_base.log
SPEEDY_K2_FIELD=1 timeout 300 python tools/bench_transforms.py 30 _k2field > gpurun_out/r2a_xf_k2f.log 2>&1; tail -5 gpurun_out/r2a_xf_k2f.log
for m in 1 8 16; do timeout 300 python tools/ktime.py $m 2>&1 | tail -2; done
for m in 8 16; do SPEEDY_K2_FIELD=1 timeout 300 python tools/ktime.py $m 2>&1 | tail -2; done
