#!/bin/bash
# Run under gpurun: per-SASS-instruction execution counts of one full-size k_trace launch (where do the issue slots go?)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-em"
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 60 -c 1 -o gpurun_out/prof_trace_sass -f $B > gpurun_out/ncu_trace_sass.log 2>&1
ncu -i gpurun_out/prof_trace_sass.ncu-rep --page source --csv --print-source sass > gpurun_out/trace_sass.csv 2>/dev/null
rm -f gpurun_out/prof_trace_sass.ncu-rep
wc -l gpurun_out/trace_sass.csv
