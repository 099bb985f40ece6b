#!/bin/bash
# full GPU test suite + sanitizer on the quad kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2w_pytest.log
timeout 600 compute-sanitizer --tool memcheck python tools/memcheck_run.py t30x8 > gpurun_out/r2w_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2w_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/memcheck_run.py t30x8 > gpurun_out/r2w_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2w_racecheck.log
