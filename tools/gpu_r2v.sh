#!/usr/bin/env bash
# Session V (1 GPU): planner rule for single-split multi-wave grids: full GPU tests, int rates, bench N=1.
set -uo pipefail
O=gpurun_out/r2v; mkdir -p $O
timeout 300 python tools/time_modes.py 1048576 int8_sim,int4_sim > $O/int_rates.log 2>&1; echo "rates rc=$?"; cat $O/int_rates.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > $O/gputests.log 2>&1; echo "pytest rc=$?"; tail -n 4 $O/gputests.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -c 300 $O/bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2v/bench_n1.json").read().strip().splitlines()[-1])
print("value %.4e ms/step %.3f kernel_ms %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"]), d["roofline"]["frac"])
print({k: round(v["value"]/1e12,3) for k,v in d["other_modes"].items()})
PY
