#!/bin/bash
# steady-state L2 behaviour of the step kernels (caches not flushed between kernels, application replay): DRAM bytes and L2 hit rate per launch
mkdir -p gpurun_out
for m in 4 8; do
timeout 500 ncu --cache-control none --replay-mode application --clock-control none -k regex:"k_grid_columns|k_spec_step|k_s2g_quad|k_g2s_quad|k_s2g_stream|k_g2s_stream" -s 200 -c 8 \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_bytes.sum --csv --log-file gpurun_out/l2_m$m.csv python tools/run_members.py $m 2 > gpurun_out/l2_m$m.log 2>&1; echo "m$m rc=$?"
done
