#!/usr/bin/env bash
# Round-2 GPU session K (1 GPU): int4 chaos spread, full GPU tests, smoke, bench N=1 (both arms; reference arm also under torchrun).
set -uo pipefail
O=gpurun_out/r2k; mkdir -p $O
timeout 300 python tools/int4_chaos.py c1_disk5000 12 > $O/int4_chaos_n5000.log 2>&1; echo "chaos rc=$?"; tail -9 $O/int4_chaos_n5000.log
timeout 200 python tools/int4_chaos.py c1_disk2000 8 > $O/int4_chaos_n2000.log 2>&1; echo "chaos rc=$?"; tail -7 $O/int4_chaos_n2000.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/gputests.log 2>&1; echo "pytest rc=$?"; tail -15 $O/gputests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
timeout 700 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -c 400 $O/bench_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $O/bench_reference_torchrun.json 2> $O/bench_reference.err; echo "reference arm rc=$?"; cat $O/bench_reference_torchrun.json | cut -c1-900
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2k/bench_n1.json").read().strip().splitlines()[-1])
print("value %.4e ms/step %.3f kernel_ms %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"]), d["roofline"]["frac"], d["roofline"]["peak"])
print(json.dumps(d.get("small_n")))
print(json.dumps(d.get("parity")))
PY
