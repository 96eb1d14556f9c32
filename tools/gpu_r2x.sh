#!/usr/bin/env bash
# Session X (1 GPU): final binary — GPU tests, smoke, bench N=1 (short).
set -uo pipefail
O=gpurun_out/r2x; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q --timeout 600 > $O/gputests.log 2>&1; echo "pytest rc=$?"; tail -n 2 $O/gputests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --no-extras > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; cut -c1-260 $O/bench_n1.json
