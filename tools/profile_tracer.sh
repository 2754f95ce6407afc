#!/bin/bash
# Run under gpurun: launch list of a short bench + one full ncu capture of the dominant kernels.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 60 -c 2 -o gpurun_out/prof_extend -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_extend.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 60 -c 1 -o gpurun_out/prof_shade -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_shade.log 2>&1
ls -la gpurun_out
