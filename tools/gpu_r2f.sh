#!/usr/bin/env bash
# Round-2 GPU session F (1 GPU): one-barrier small-N kernel (guarded), fp32 schedule variants, full tests, bench.
set -uo pipefail
O=gpurun_out/r2f; mkdir -p $O
timeout 180 python -m pytest tests/test_gpu_persistent.py -x -q --timeout 120 > $O/persistent_tests.log 2>&1; echo "persistent tests rc=$?"; tail -15 $O/persistent_tests.log
NB_B200_ONE_BARRIER=1 timeout 120 python tools/time_small.py > $O/small_one_barrier.log 2>&1; echo "rc=$?"; cat $O/small_one_barrier.log
NB_B200_ONE_BARRIER=0 timeout 120 python tools/time_small.py > $O/small_two_barriers.log 2>&1; echo "rc=$?"; cat $O/small_two_barriers.log
for b in tools/variants_f32/v_*; do timeout 60 $b >> $O/f32_variants.log 2>&1; done; sort -k9 -n $O/f32_variants.log | head -40
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > $O/gputests.log 2>&1; echo "pytest rc=$?"; tail -12 $O/gputests.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -c 400 $O/bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2f/bench_n1.json").read().strip().splitlines()[-1])
print("value %.4e ms/step %.3f kernel_ms %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"]), d["roofline"]["frac"], d["roofline"]["peak"])
print(json.dumps(d["small_n"]))
PY
