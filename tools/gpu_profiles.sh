#!/bin/bash
# round-2 evidence: bench line, ncu launch list of the same command, ncu --set full of the step kernels (1 and 8 members), transform micro-benchmark
tag=${1:-r2}
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file gpurun_out/${tag}_ncu_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ensemble > gpurun_out/${tag}_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_s2g_stream|k_grid_columns|k_g2s_stream|k_spec_step" -s 200 -c 4 -f -o gpurun_out/${tag}_prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ensemble > gpurun_out/${tag}_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_grid_columns|k_spec_step|k_s2g_quad|k_g2s_quad" -s 40 -c 4 -f -o gpurun_out/${tag}_prof_m8 python tools/run_members.py 8 1 > gpurun_out/${tag}_ncu3.log 2>&1; echo "ncu3 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/${tag}_ncu_launches_m8.csv python tools/run_members.py 8 1 > gpurun_out/${tag}_ncu4.log 2>&1; echo "ncu4 rc=$?"
timeout 300 python tools/bench_transforms.py 30 _${tag} > gpurun_out/${tag}_xf.log 2>&1; tail -10 gpurun_out/${tag}_xf.log | cut -c1-140
timeout 300 python tools/configs_bench.py > gpurun_out/${tag}_configs.json 2> gpurun_out/${tag}_configs.err; cat gpurun_out/${tag}_configs.json | cut -c1-400
for m in 1 8 16; do timeout 200 python tools/ktime.py $m 2>&1 | tail -2 | cut -c1-260; done
