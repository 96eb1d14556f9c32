#!/usr/bin/env bash
# Session T (1 GPU): ncu --set full of the fast-lookup (INT8_SIM) pair kernel of the FINAL build at N = 2^20.
set -uo pipefail
O=gpurun_out/r2t; mkdir -p $O
timeout 200 python tools/prof_force.py 1048576 int8_sim f32 > $O/prof_plain.log 2>&1; echo "prof plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:accel_kernel --launch-skip 1 --launch-count 1 \
    -o $O/force_int8_sim_n1m -f python tools/prof_force.py 1048576 int8_sim f32 > $O/ncu_full_int8_sim.log 2>&1; echo "ncu full rc=$?"
ls -la $O
