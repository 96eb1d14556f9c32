#!/usr/bin/env bash
# Round-2 GPU session A: full GPU test suite, bench (N=1), int-kernel launch-config variants, ncu launch list + full captures.
set -uo pipefail
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/gputests.log 2>&1; echo "pytest rc=$?"; tail -15 $O/gputests.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -c 1500 $O/bench_n1.err; head -c 3000 $O/bench_n1.json
./tools/peaks > $O/pipe_peaks.json 2>&1
for v in base a b c d e f g h; do
  echo "== variant $v" >> $O/variants.log
  NB_B200_LIB=tools/variants/libnb_$v.so timeout 300 python tools/time_modes.py 131072 int8_sim,int4_sim >> $O/variants.log 2>&1
done
tail -40 $O/variants.log
# launch list of the bench command (after it exited 0 without ncu above)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
# full captures of the final kernels at the benchmark N
for spec in "float32 f32" "float64 f64" "int8_sim f32"; do
  set -- $spec
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:accel_kernel --launch-skip 1 --launch-count 1 \
      -o $O/force_${1}_n1m -f python tools/prof_force.py 1048576 $1 $2 > $O/ncu_full_$1.log 2>&1; echo "ncu full $1 rc=$?"
done
ls -la $O
