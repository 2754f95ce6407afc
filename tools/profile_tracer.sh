#!/bin/bash
# Run under gpurun: ncu launch list of a short bench + full captures of the dominant kernels (never a bench value).
# The .ncu-rep files are summarised on the box and removed: gpurun merges at most 64 MiB back.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-em --no-strong --no-config5"
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/bench_under_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/launches.csv > gpurun_out/launches_tracer.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 60 -c 2 -o gpurun_out/prof_trace -f $B > gpurun_out/ncu_trace.log 2>&1
python tools/ncu_summary.py rep gpurun_out/prof_trace.ncu-rep > gpurun_out/ncu_trace.txt 2>&1
ncu -i gpurun_out/prof_trace.ncu-rep --page details --csv 2>/dev/null | grep -E "Pipe|pipe|Stall|stall|Issue|Eligible|Active Warps|Achieved Occupancy|Branch" | head -80 > gpurun_out/ncu_trace_details.txt
rm -f gpurun_out/prof_trace.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 60 -c 1 -o gpurun_out/prof_shade -f $B > gpurun_out/ncu_shade.log 2>&1
python tools/ncu_summary.py rep gpurun_out/prof_shade.ncu-rep > gpurun_out/ncu_shade.txt 2>&1
rm -f gpurun_out/prof_shade.ncu-rep
ncu --set full --clock-control none -k regex:k_guiding_update -c 2 -o gpurun_out/prof_guiding -f python tools/guiding_time.py 57600 > gpurun_out/ncu_guiding.log 2>&1
python tools/ncu_summary.py rep gpurun_out/prof_guiding.ncu-rep > gpurun_out/ncu_guiding.txt 2>&1
ncu -i gpurun_out/prof_guiding.ncu-rep --page details --csv 2>/dev/null | grep -E "Pipe|pipe|Stall|stall|Issue|Eligible|Active Warps|Achieved Occupancy|Branch" | head -80 > gpurun_out/ncu_guiding_details.txt
rm -f gpurun_out/prof_guiding.ncu-rep
B200PT_GUIDING_ORDER=strict ncu --set full --clock-control none -k regex:k_guiding_update -c 1 -o gpurun_out/prof_guiding_strict -f python tools/guiding_time.py 57600 > gpurun_out/ncu_guiding_strict.log 2>&1
python tools/ncu_summary.py rep gpurun_out/prof_guiding_strict.ncu-rep > gpurun_out/ncu_guiding_strict.txt 2>&1
rm -f gpurun_out/prof_guiding_strict.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_guiding.csv python tools/guiding_time.py 57600 > gpurun_out/guiding_under_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/launches_guiding.csv > gpurun_out/launches_guiding.txt 2>&1
rm -f gpurun_out/launches.csv gpurun_out/launches_guiding.csv
du -sh gpurun_out; cat gpurun_out/launches_tracer.txt; cat gpurun_out/ncu_trace.txt | head -24; cat gpurun_out/launches_guiding.txt; head -24 gpurun_out/ncu_guiding.txt
