#!/bin/bash
# Run under gpurun: per-SASS-instruction execution counts of one full-size launch of kernel $1 (regex), skipping $2 launches
# usage: tools/profile_kernel_sass.sh k_shade 60   ->  gpurun_out/sass_k_shade.csv (+ details_k_shade.txt, summary ncu_k_shade.txt)
#        PROF_CMD="python tools/run_config5.py 1920 1080 1 2" PROF_TAG=guided tools/profile_kernel_sass.sh k_shade 5
K=${1:-k_trace}; S=${2:-60}
TAG=${PROF_TAG:-$K}
mkdir -p gpurun_out
B=${PROF_CMD:-"python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-em"}
ncu --set full --clock-control none --import-source on -k "regex:$K" -s $S -c 1 -o gpurun_out/prof_sass_$TAG -f $B > gpurun_out/ncu_sass_$TAG.log 2>&1
ncu -i gpurun_out/prof_sass_$TAG.ncu-rep --page source --csv --print-source sass > gpurun_out/sass_$TAG.csv 2>/dev/null
ncu -i gpurun_out/prof_sass_$TAG.ncu-rep --page details > gpurun_out/details_$TAG.txt 2>/dev/null
python tools/ncu_summary.py rep gpurun_out/prof_sass_$TAG.ncu-rep > gpurun_out/ncu_$TAG.txt 2>&1
rm -f gpurun_out/prof_sass_$TAG.ncu-rep
wc -l gpurun_out/sass_$TAG.csv; cat gpurun_out/ncu_$TAG.txt
