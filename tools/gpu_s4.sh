#!/bin/bash
# K1 quad: probe, variant tests, micro-benchmark on/off, step timing
tag=${1:-r2x}
mkdir -p gpurun_out
timeout 60 python tools/k1q_probe.py 584 > gpurun_out/${tag}_probe.log 2>&1 || { echo PROBE FAILED; tail -5 gpurun_out/${tag}_probe.log; exit 0; }
cat gpurun_out/${tag}_probe.log
timeout 60 python tools/k1q_probe.py 1203 2>&1 | tail -4
timeout 600 python -m pytest tests/test_kernel_variants_gpu.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
for v in 0 1; do SPEEDY_K1_QUAD=$v timeout 200 python tools/bench_transforms.py 30 _k1q$v 2>&1 | grep spec_to_grid | tail -2 | cut -c1-120; done
for m in 8 16; do timeout 200 python tools/ktime.py $m 2>&1 | tail -2; done
