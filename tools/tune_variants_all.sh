#!/bin/bash
# Run under gpurun: A/B of library builds in gpurun_scratch/lib_*.so (B200PT_LIB override) on configs 2 (bench), 4 and 5
mkdir -p gpurun_out; out=gpurun_out/variants_all.txt; : > $out
for lib in gpurun_scratch/lib_*.so; do
  export B200PT_LIB=$PWD/$lib
  line=$(python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-em 2>&1 | tail -1)
  echo "$(basename $lib) config2 $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("Mrays/s=%.1f trace_Mrays/s=%.1f trace_ms=%.2f shade_ms=%.2f dev_ms=%.2f" % (d["value"], d["roofline"]["trace_Mrays_per_s"], d["stage_ms"]["trace"], d["stage_split"]["shade"], d["stage_ms"]["device"]))' 2>&1 | tail -1)" >> $out
  python tools/run_config4.py 1920 1080 4 4 2>/dev/null | python -c '
import sys, json
for l in sys.stdin:
    d = json.loads(l); print("   config4", d["run"], " ".join("%s: dev %.0f ms trace %.0f shade+ic %.0f (%.0f Mrays/s)" % (p["phase"], p["device_ms"], p["trace_ms"], p["shade_and_ic_ms"], p["Mrays_per_s"]) for p in d["phases"]), "mean %.6f" % d["image_mean"])' >> $out
  python tools/run_config5.py 2>/dev/null | python -c '
import sys, json
d = json.loads(sys.stdin.read()); print("   config5", " | ".join("%s: render %.0f ms fit %.0f ms (%.0f Mrays/s)" % (p["phase"], p["render_ms"], p["guiding_fit_ms"], p["Mrays_per_s"]) for p in d["phases"]), "mean %.6f" % d["image_mean"])' >> $out
done
cat $out
