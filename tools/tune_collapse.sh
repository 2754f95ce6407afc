#!/bin/bash
# Run under gpurun: BVH8 collapse (DP vs greedy) and the DP's node cost
mkdir -p gpurun_out; out=gpurun_out/tune_collapse.txt; : > $out
run() { line=$(env "$@" python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-em 2>gpurun_out/err.txt | tail -1)
  echo "$* $(grep bvh8 gpurun_out/err.txt | head -1 | cut -c1-140)" >> $out
  echo "   $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("Mrays/s=%.1f e2e=%.1f trace_Mrays/s=%.1f trace_ms=%.2f dev_ms=%.2f" % (d["value"], d["e2e"]["value"], d["roofline"]["trace_Mrays_per_s"], d["stage_ms"]["trace"], d["stage_ms"]["device"]))' 2>&1 | tail -1)" >> $out; }
export B200PT_BVH_STATS=1
run B200PT_BVH_GREEDY=1
run B200PT_BVH_NODE_COST=1.0
run B200PT_BVH_NODE_COST=2.0
run B200PT_BVH_NODE_COST=3.0
run B200PT_BVH_NODE_COST=5.0
run B200PT_BVH_NODE_COST=2.0 B200PT_BVH_TRAV_COST=1.0
run B200PT_BVH_NODE_COST=2.0 B200PT_BVH_TRAV_COST=0.25
cat $out
