#!/bin/bash
# Run under gpurun: A/B of library builds placed in gpurun_scratch/lib_*.so (B200PT_LIB override)
mkdir -p gpurun_out; out=gpurun_out/variants.txt; : > $out
for lib in gpurun_scratch/lib_*.so; do
  line=$(B200PT_LIB=$PWD/$lib python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-em 2>&1 | tail -1)
  echo "$(basename $lib) $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("Mrays/s=%.1f trace_Mrays/s=%.1f trace_ms=%.2f shade_ms=%.2f dev_ms=%.2f" % (d["value"], d["roofline"]["trace_Mrays_per_s"], d["stage_ms"]["trace"], d["stage_split"]["shade"], d["stage_ms"]["device"]))' 2>&1 | tail -1)" >> $out
done
cat $out
