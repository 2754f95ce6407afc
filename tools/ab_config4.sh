#!/bin/bash
# Run under gpurun: A/B of the IC-frame knobs on config 4; images must be identical
# usage: tools/ab_config4.sh "ENV=.. ENV=.." "ENV=.." ...   (one run per argument; default: the three knobs off / on)
mkdir -p gpurun_out
OUT=gpurun_out/ab_config4.txt
: > $OUT
if [ $# -eq 0 ]; then set -- "B200PT_ICQ_SORT=0 B200PT_REGEN_SPLIT=0" "B200PT_SHADE_SORTED=0" "B200PT_SHADE_SORTED=1"; fi
for combo in "$@"; do
  echo "# $combo (no stage events)" >> $OUT
  env $combo RUN4_STAGE=0 RUN4_RUNS=${RUN4_RUNS:-ic,adrrs} python tools/run_config4.py >> $OUT 2>&1
done
python - <<'PY' > gpurun_out/ab_config4_summary.txt
import json
for l in open("gpurun_out/ab_config4.txt"):
    if l.startswith("#"): print(l.strip()); continue
    try: d = json.loads(l)
    except Exception: continue
    print("   %-12s mean %.10f  " % (d["run"], d["image_mean"]) + "  ".join("%s %.1f ms (%.0f Mrays/s)" % (p["phase"], p["device_ms"], p["Mrays_per_s"]) for p in d["phases"]))
PY
cat gpurun_out/ab_config4_summary.txt
