#!/usr/bin/env bash
# Session O (1 GPU): ncu evidence of the FINAL build: launch list of the bench command, --set full of the fp32 pair kernel at N = 2^20.
set -uo pipefail
O=gpurun_out/r2o; mkdir -p $O
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_plain.json 2> $O/bench_plain.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 120 python tools/prof_force.py 1048576 float32 f32 > $O/prof_plain.log 2>&1; echo "prof plain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:accel_kernel --launch-skip 1 --launch-count 1 \
    -o $O/force_float32_n1m -f python tools/prof_force.py 1048576 float32 f32 > $O/ncu_full_float32.log 2>&1; echo "ncu full rc=$?"
ls -la $O
