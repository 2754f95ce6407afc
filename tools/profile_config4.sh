#!/bin/bash
# Run under gpurun: config 4 (sponzaXML 1080p, IC_SIZE 10000) with per-stage CUDA-event times, then an ncu launch list
# of a short IC run (kernel shares of the step; never a bench value).
mkdir -p gpurun_out
python tools/run_config4.py 1920 1080 4 4 > gpurun_out/config4_stage.jsonl 2> gpurun_out/config4_stage.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/config4_launches.csv \
    python tools/run_config4_short.py > gpurun_out/config4_under_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/config4_launches.csv > gpurun_out/config4_launch_summary.txt 2>&1 || true
rm -f gpurun_out/config4_launches.csv; cat gpurun_out/config4_stage.jsonl; tail -30 gpurun_out/config4_launch_summary.txt
