mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:^k_ic_create$ -s 0 -c 2 -o gpurun_out/prof_ic_create -f python tools/run_config4_short.py > gpurun_out/ncu_ic_create.log 2>&1
python tools/ncu_summary.py rep gpurun_out/prof_ic_create.ncu-rep > gpurun_out/ncu_ic_create.txt 2>&1
ncu -i gpurun_out/prof_ic_create.ncu-rep --page source --csv --print-source sass > gpurun_out/sass_k_ic_create.csv 2>/dev/null
rm -f gpurun_out/prof_ic_create.ncu-rep
cat gpurun_out/ncu_ic_create.txt
