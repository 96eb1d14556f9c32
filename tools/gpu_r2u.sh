#!/usr/bin/env bash
# Session U (1 GPU): j-split count of the fast-lookup kernel at N = 2^20 (the planner's wave model is flat in s there and keeps 1 split).
set -uo pipefail
O=gpurun_out/r2u; mkdir -p $O
for s in 0 4 10; do
  echo "--- NB_B200_SPLITS=$s" >> $O/int_splits.log
  NB_B200_SPLITS=$s timeout 200 python tools/time_modes.py 1048576 int8_sim,float16 >> $O/int_splits.log 2>&1
done
echo "--- NB_B200_SPLITS=13 (cap lifted)" >> $O/int_splits.log
NB_B200_SPLITS=13 NB_B200_SPLIT_WORKSPACE_MB=2048 NB_B200_SPLIT_CAP=64 timeout 200 python tools/time_modes.py 1048576 int8_sim,float16 >> $O/int_splits.log 2>&1
cat $O/int_splits.log
