#!/bin/bash
# Run under gpurun: config-4 short run (6 IC prepare frames, estimate frame, 2 ADRRS+split frames): ncu launch list, then a
# full capture with source counters of one heavy k_shade<.,IC> launch ($1 = launches to skip) -> per-line table.
S=${1:-260}
mkdir -p gpurun_out
B="python tools/run_config4_short.py"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/config4_launches.csv $B > gpurun_out/config4_under_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/config4_launches.csv > gpurun_out/config4_launches.txt 2>&1
grep -n "k_shade" gpurun_out/config4_launches.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | tail -3
rm -f gpurun_out/config4_launches.csv
ncu --set full --clock-control none --import-source on -k regex:k_shade -s $S -c 1 -o gpurun_out/prof_shade_ic -f $B > gpurun_out/ncu_shade_ic.log 2>&1
python tools/ncu_summary.py rep gpurun_out/prof_shade_ic.ncu-rep > gpurun_out/ncu_shade_ic.txt 2>&1
ncu -i gpurun_out/prof_shade_ic.ncu-rep --page source --csv --print-source sass > gpurun_out/sass_k_shade_ic.csv 2>/dev/null
ncu -i gpurun_out/prof_shade_ic.ncu-rep --page details > gpurun_out/details_k_shade_ic.txt 2>/dev/null
rm -f gpurun_out/prof_shade_ic.ncu-rep
cat gpurun_out/config4_launches.txt | head -12; cat gpurun_out/ncu_shade_ic.txt
