#!/bin/bash
tag=${1:-r2x}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
for m in 1 8 16; do timeout 200 python tools/ktime.py $m 2>&1 | tail -2 | cut -c1-400 | tr '\n' ' '; echo; done
timeout 200 python tools/ktime.py 8 30 sppt 2>&1 | tail -1 | cut -c1-100
m=8
timeout 500 ncu --cache-control none --replay-mode application --clock-control none -k regex:"k_grid_columns|k_spec_step|k_s2g_quad|k_g2s_quad" -s 200 -c 4 \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_bytes.sum --csv --log-file gpurun_out/${tag}_l2_m$m.csv python tools/run_members.py $m 2 > gpurun_out/${tag}_l2_m$m.log 2>&1; echo "l2 m$m rc=$?"
