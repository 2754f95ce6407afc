#!/bin/bash
# Run under gpurun: light per-launch metrics of every k_shade launch of the short config-4 run, then a full capture of a heavy one
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:k_shade --csv --log-file gpurun_out/shade_ic_launches.csv python tools/run_config4_short.py > gpurun_out/ncu_shade_ic.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.DictReader(l for l in open("gpurun_out/shade_ic_launches.csv") if not l.startswith("=="))]
by={}
for r in rows: by.setdefault(int(r["ID"]),{})[r["Metric Name"]]=float(r["Metric Value"].replace(",",""))
ids=sorted(by)
print("launches",len(ids))
heavy=sorted(ids,key=lambda i:-by[i]["gpu__time_duration.sum"])[:12]
for n,i in enumerate(ids):
    if i in heavy or n%60==0: print(n,i,by[i])
open("gpurun_out/heavy_index.txt","w").write(str(ids.index(heavy[0])))
PY
SKIP=$(cat gpurun_out/heavy_index.txt)
ncu --set full --clock-control none --import-source on -k regex:k_shade -s $SKIP -c 1 -o gpurun_out/prof_shade_ic -f python tools/run_config4_short.py >> gpurun_out/ncu_shade_ic.log 2>&1
python tools/ncu_summary.py rep gpurun_out/prof_shade_ic.ncu-rep > gpurun_out/prof_shade_ic.txt 2>&1
cat gpurun_out/prof_shade_ic.txt
