#!/bin/bash
# Run under gpurun: knobs of the cooperative-triangle trace kernel (refill threshold, chunk, CTAs/SM) and the BVH leaf policy
mkdir -p gpurun_out; out=gpurun_out/tune_coop.txt; : > $out
run() { line=$(env "$@" python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-em 2>&1 | tail -1)
  echo "$* $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("Mrays/s=%.1f trace_Mrays/s=%.1f trace_ms=%.2f shade_ms=%.2f frame_ms=%.2f dev_ms=%.2f" % (d["value"], d["roofline"]["trace_Mrays_per_s"], d["stage_ms"]["trace"], d["stage_split"]["shade"], d["stage_ms"]["frame_total"], d["stage_ms"]["device"]))' 2>&1 | tail -1)" >> $out; }
run X=1
run B200PT_TRACE_REFILL=4
run B200PT_TRACE_REFILL=12
run B200PT_TRACE_REFILL=16
run B200PT_TRACE_REFILL=24
run B200PT_TRACE_CHUNK=32
run B200PT_TRACE_CHUNK=128
run B200PT_TRACE_GRID_PER_SM=6
run B200PT_TRACE_GRID_PER_SM=5
run B200PT_BVH_MAX_LEAF=2
run B200PT_BVH_MAX_LEAF=1
run B200PT_BVH_TRAV_COST=0.25
run B200PT_BVH_TRAV_COST=1.0
run B200PT_BVH_TRAV_COST=2.0
cat $out
