#!/usr/bin/env bash
# Session P (4 GPUs): the world size not exercised so far — smoke + bench --gpus 4.
set -uo pipefail
O=gpurun_out/r2p; mkdir -p $O
W=${NB_WORLD:-4}
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29561 \
    tools/sharded_smoke.py > $O/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; grep "SMOKE\|Error" $O/smoke.log | tail -n 4
NB_BENCH_WATCHDOG_S=200 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29562 \
    bench.py --gpus $W --steps 5 --warmup 3 > $O/bench_n$W.json 2> $O/bench_n$W.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_n$W.json").read().strip().splitlines()[-1])
    print("value %.4e ms/step %.3f e2e %.4e kernel_ms %.3f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["gpu_launches"]))
    print("  parity", (d.get("parity") or {}).get("status"), (d.get("parity") or {}).get("bit_identical_to_world1"))
    for k,v in (d.get("lines") or {}).items(): print("  ", k, "%.4e" % v["value"], "ms %.2f" % v["ms_per_step"])
except Exception as e:
    print("ERR", e)
PY
