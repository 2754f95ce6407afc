#!/bin/bash
# Run under gpurun: ncu launch list of the ADRRS frames of config 4 (50 prepare frames + estimate skipped in the summary)
mkdir -p gpurun_out
RUN4_PREPARE=50 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/config4_adrrs_launches.csv \
    python tools/run_config4_short.py > gpurun_out/config4_adrrs_under_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/config4_adrrs_launches.csv 51 > gpurun_out/config4_adrrs_launch_summary.txt 2>&1
python tools/ncu_summary.py launches gpurun_out/config4_adrrs_launches.csv 0 > gpurun_out/config4_all_launch_summary.txt 2>&1
rm -f gpurun_out/config4_adrrs_launches.csv
cat gpurun_out/config4_adrrs_launch_summary.txt; tail -5 gpurun_out/config4_adrrs_under_ncu.log
