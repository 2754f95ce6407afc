#!/bin/bash
# Run under gpurun: ncu launch list of a short bench + full captures of the dominant kernels (never a bench value).
set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-em"
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 60 -c 2 -o gpurun_out/prof_trace -f $B > gpurun_out/ncu_trace.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 60 -c 1 -o gpurun_out/prof_shade -f $B > gpurun_out/ncu_shade.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_guiding_update -c 2 -o gpurun_out/prof_guiding -f python tools/guiding_time.py 57600 > gpurun_out/ncu_guiding.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_guiding.csv python tools/guiding_time.py 57600 > gpurun_out/guiding_under_ncu.log 2>&1
ls -la gpurun_out
