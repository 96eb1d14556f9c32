#!/usr/bin/env bash
# Round-2 GPU session C (1 GPU): persistent small-N kernel (guarded by timeouts), full GPU tests, split-count A/B.
set -uo pipefail
O=gpurun_out/r2c; mkdir -p $O
timeout 180 python -m pytest tests/test_gpu_persistent.py -x -q --timeout 120 > $O/persistent_tests.log 2>&1; echo "persistent tests rc=$?"; tail -15 $O/persistent_tests.log
nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader
NB_B200_PERSISTENT=1 timeout 120 python tools/time_small.py > $O/small_on.log 2>&1; echo "rc=$?"; cat $O/small_on.log
NB_B200_PERSISTENT=0 timeout 120 python tools/time_small.py > $O/small_off.log 2>&1; echo "rc=$?"; cat $O/small_off.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > $O/gputests.log 2>&1; echo "pytest rc=$?"; tail -12 $O/gputests.log
for s in 3 8 3 8 6 10; do NB_B200_SPLITS=$s timeout 200 python tools/time_splits.py >> $O/splits.log 2>&1; done; cat $O/splits.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -c 400 $O/bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c/bench_n1.json").read().strip().splitlines()[-1])
print("value %.4e ms/step %.3f kernel_ms %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"]), d["roofline"]["frac"], d["roofline"]["peak"])
print(json.dumps(d["small_n"]))
print(json.dumps(d["potential_energy_fused"]))
PY
