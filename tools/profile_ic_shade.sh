#!/bin/bash
# Run under gpurun: full ncu capture of a heavy k_shade<.,IC> launch and a k_ic_query launch of the short config-4 run
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:k_shade -s 210 -c 1 -o gpurun_out/prof_shade_ic -f python tools/run_config4_short.py > gpurun_out/ncu_shade_ic.log 2>&1
python tools/ncu_summary.py rep gpurun_out/prof_shade_ic.ncu-rep > gpurun_out/ncu_shade_ic.txt 2>&1
rm -f gpurun_out/prof_shade_ic.ncu-rep
ncu --set full --clock-control none -k regex:k_ic_query -s 210 -c 1 -o gpurun_out/prof_ic_query -f python tools/run_config4_short.py >> gpurun_out/ncu_shade_ic.log 2>&1
python tools/ncu_summary.py rep gpurun_out/prof_ic_query.ncu-rep > gpurun_out/ncu_ic_query.txt 2>&1
rm -f gpurun_out/prof_ic_query.ncu-rep
cat gpurun_out/ncu_shade_ic.txt gpurun_out/ncu_ic_query.txt
