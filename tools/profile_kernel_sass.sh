#!/bin/bash
# Run under gpurun: per-SASS-instruction execution counts of one full-size launch of kernel $1 (regex), skipping $2 launches
# usage: tools/profile_kernel_sass.sh k_shade 60   ->  gpurun_out/sass_k_shade.csv (+ source-level view sass_k_shade_src.csv)
K=${1:-k_trace}; S=${2:-60}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-em"
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -o gpurun_out/prof_sass_$K -f $B > gpurun_out/ncu_sass_$K.log 2>&1
ncu -i gpurun_out/prof_sass_$K.ncu-rep --page source --csv --print-source sass > gpurun_out/sass_$K.csv 2>/dev/null
ncu -i gpurun_out/prof_sass_$K.ncu-rep --page source --csv --print-source cuda > gpurun_out/sass_${K}_src.csv 2>/dev/null
ncu -i gpurun_out/prof_sass_$K.ncu-rep --page details > gpurun_out/details_$K.txt 2>/dev/null
rm -f gpurun_out/prof_sass_$K.ncu-rep
wc -l gpurun_out/sass_$K.csv gpurun_out/sass_${K}_src.csv
