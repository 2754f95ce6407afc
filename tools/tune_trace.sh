#!/bin/bash
# Run under gpurun: sweep of the persistent-trace tuning knobs (env overrides read in b200pt_create).
mkdir -p gpurun_out
out=gpurun_out/tune.txt
: > $out
for cap in 1 2 3 4 6 24; do for refill in 4 8; do
  line=$(B200PT_TRACE_TRICAP=$cap B200PT_TRACE_REFILL=$refill python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-em 2>&1 | tail -1)
  echo "tricap=$cap refill=$refill $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("Mrays/s=%.1f trace_Mrays/s=%.1f trace_ms=%.2f shade_ms=%.2f dev_ms=%.2f" % (d["value"], d["roofline"]["trace_Mrays_per_s"], d["stage_ms"]["trace"], d["stage_split"]["shade"], d["stage_ms"]["device"]))' 2>&1 | tail -1)" >> $out
done; done
cat $out
