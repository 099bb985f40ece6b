#!/bin/bash
tag=${1:-r2x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernel_variants_gpu.py tests/test_ensemble_gpu.py tests/test_output_gpu.py -x -q -k "48h or ensemble or partition or small_blocks or restart or sppt or plumbing or plumbing" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
for v in 1 0; do for m in 6 8 12 16; do echo -n "alias=$v m$m: "; SPEEDY_TRANSIENT_ALIAS=$v timeout 200 python tools/ktime.py $m 2>&1 | tail -2 | cut -c1-400 | tr '\n' ' '; echo; done; done
for v in 1 0; do echo -n "alias=$v sppt m8: "; SPEEDY_TRANSIENT_ALIAS=$v timeout 200 python tools/ktime.py 8 30 sppt 2>&1 | tail -1 | cut -c1-100; done
m=8
timeout 500 ncu --cache-control none --replay-mode application --clock-control none -k regex:"k_grid_columns|k_spec_step|k_s2g_quad|k_g2s_quad" -s 200 -c 4 \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_bytes.sum --csv --log-file gpurun_out/${tag}_l2_m$m.csv python tools/run_members.py $m 2 > gpurun_out/${tag}_l2_m$m.log 2>&1; echo "l2 m$m rc=$?"
