#!/usr/bin/env bash
# Session N (1 GPU): int4 chaos with the all-coordinate ulp perturbation, full GPU tests.
set -uo pipefail
O=gpurun_out/r2n; mkdir -p $O
timeout 300 python tools/int4_chaos.py c1_disk5000 16 all > $O/int4_chaos_n5000_all.log 2>&1; echo "chaos rc=$?"; tail -n 12 $O/int4_chaos_n5000_all.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/gputests.log 2>&1; echo "pytest rc=$?"; tail -n 6 $O/gputests.log
