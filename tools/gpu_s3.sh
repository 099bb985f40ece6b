#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/k2f_probe.py k2_quad 584 > gpurun_out/r2c_probe.log 2>&1 || { echo PROBE FAILED; tail -5 gpurun_out/r2c_probe.log; timeout 600 compute-sanitizer --tool memcheck python tools/k2f_probe.py k2_quad 584 2>&1 | grep -v "^$" | head -60 > gpurun_out/r2c_memcheck.log; head -40 gpurun_out/r2c_memcheck.log; exit 0; }
cat gpurun_out/r2c_probe.log
timeout 900 python -m pytest tests/test_kernel_variants_gpu.py -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2c_pytest.log
SPEEDY_K2_QUAD=1 timeout 300 python tools/bench_transforms.py 30 _k2quad > gpurun_out/r2c_xf.log 2>&1; tail -5 gpurun_out/r2c_xf.log
for m in 8 16; do SPEEDY_K2_QUAD=1 timeout 300 python tools/ktime.py $m 2>&1 | tail -2; done
