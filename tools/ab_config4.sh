#!/bin/bash
# Run under gpurun: A/B of the IC-frame knobs on config 4 (cell-sorted lookups, regen kernel), images must be identical
mkdir -p gpurun_out
OUT=gpurun_out/ab_config4.txt
: > $OUT
for combo in "0 0" "1 0" "0 1" "1 1"; do
  set -- $combo
  echo "# B200PT_ICQ_SORT=$1 B200PT_REGEN_SPLIT=$2 (no stage events)" >> $OUT
  B200PT_ICQ_SORT=$1 B200PT_REGEN_SPLIT=$2 RUN4_STAGE=0 RUN4_RUNS=ic,adrrs python tools/run_config4.py >> $OUT 2>&1
done
echo "# defaults, with stage events" >> $OUT
RUN4_RUNS=ic,adrrs python tools/run_config4.py >> $OUT 2>&1
cat $OUT
