#!/bin/bash
# Run under gpurun: leaf-size policy of the BVH builder (host side; hits are identical, tests check that) and the counter path
mkdir -p gpurun_out; out=gpurun_out/tune_bvh.txt; : > $out
run() { line=$(env "$@" python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-em 2>&1 | tail -1)
  echo "$* $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("Mrays/s=%.1f trace_Mrays/s=%.1f trace_ms=%.2f shade_ms=%.2f frame_ms=%.2f dev_ms=%.2f" % (d["value"], d["roofline"]["trace_Mrays_per_s"], d["stage_ms"]["trace"], d["stage_split"]["shade"], d["stage_ms"]["frame_total"], d["stage_ms"]["device"]))' 2>&1 | tail -1)" >> $out; }
run X=1
run B200PT_COUNTER_COPY=1
run B200PT_BVH_TRAV_COST=0.25
run B200PT_BVH_TRAV_COST=0.1
run B200PT_BVH_TRAV_COST=0.0
run B200PT_BVH_MAX_LEAF=2
run B200PT_BVH_MAX_LEAF=1
run B200PT_BVH_TRAV_COST=1.0
cat $out
