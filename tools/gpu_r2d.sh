#!/usr/bin/env bash
# Round-2 GPU session D (1 GPU): full GPU tests; A/B of the two ptxas schedules of the fp32 force loop on one box.
set -uo pipefail
O=gpurun_out/r2d; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > $O/gputests.log 2>&1; echo "pytest rc=$?"; tail -12 $O/gputests.log
for w in 0 1 0 1; do NB_B200_WINDOW_KERNEL=$w timeout 200 python tools/time_splits.py >> $O/window_kernel_ab.log 2>&1; echo "WINDOW_KERNEL=$w" >> $O/window_kernel_ab.log; done; cat $O/window_kernel_ab.log
