#!/bin/bash
# usage: gpu_s3.sh <tag> <option>   — probe (short timeout), variant tests, transform micro-benchmark, 8/16-member step timing
tag=${1:-r2x}; opt=${2:-k2_quad}; OPT=$(echo $opt | tr a-z A-Z)
mkdir -p gpurun_out
timeout 90 python tools/k2f_probe.py $opt 584 > gpurun_out/${tag}_probe.log 2>&1 || { echo PROBE FAILED; tail -5 gpurun_out/${tag}_probe.log; timeout 150 compute-sanitizer --tool memcheck python tools/k2f_probe.py $opt 584 2>&1 | grep -v "^$" | head -60 > gpurun_out/${tag}_memcheck.log; head -40 gpurun_out/${tag}_memcheck.log; exit 0; }
cat gpurun_out/${tag}_probe.log
timeout 600 python -m pytest tests/test_kernel_variants_gpu.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
env SPEEDY_$OPT=1 timeout 200 python tools/bench_transforms.py 30 _$opt > gpurun_out/${tag}_xf.log 2>&1; tail -5 gpurun_out/${tag}_xf.log
for m in 8 16; do env SPEEDY_$OPT=1 timeout 200 python tools/ktime.py $m 2>&1 | tail -2; done
