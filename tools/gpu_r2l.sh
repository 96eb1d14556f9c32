#!/usr/bin/env bash
# Session L (1 GPU): split-count A/B with the workspace cap lifted, on the one-GPU shape and the 8-way shard shape.
set -uo pipefail
O=gpurun_out/r2l; mkdir -p $O
run() { env "$@" timeout 120 python tools/time_shapes.py 1048576 1 8 >> $O/splits_shapes.log 2>&1; }
run NB_B200_SPLITS=0
for s in 8 13 26; do run NB_B200_SPLITS=$s NB_B200_SPLIT_WORKSPACE_MB=2048 NB_B200_SPLIT_CAP=64; done
for s in 32 45 52 64; do env NB_B200_SPLITS=$s NB_B200_SPLIT_WORKSPACE_MB=2048 NB_B200_SPLIT_CAP=64 timeout 120 python tools/time_shapes.py 1048576 8 >> $O/splits_shapes.log 2>&1; done
run NB_B200_SPLITS=0 NB_B200_SPLIT_WORKSPACE_MB=2048 NB_B200_SPLIT_CAP=64
grep "n_tgt" $O/splits_shapes.log
