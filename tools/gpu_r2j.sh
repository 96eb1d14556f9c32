#!/usr/bin/env bash
# Session J (2 GPUs): A/B of the gather/pair-kernel arrangements, then bench N=2 with the default.
set -uo pipefail
O=gpurun_out/r2j; mkdir -p $O
W=${NB_WORLD:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29531 \
    tools/time_sharded.py > $O/overlap_ab_n$W.log 2>&1; echo "ab rc=$?"; grep "world=\|Error" $O/overlap_ab_n$W.log | tail -12
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29532 \
    tools/sharded_smoke.py > $O/smoke.log 2>&1; echo "smoke rc=$?"; grep "SMOKE\|Error" $O/smoke.log | tail -4
NB_BENCH_WATCHDOG_S=250 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $W --steps 5 --warmup 3 --no-extras > $O/bench_n$W.json 2> $O/bench_n$W.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_n$W.json").read().strip().splitlines()[-1])
    print("value %.4e ms/step %.3f e2e %.4e kernel_ms %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"]))
    print("  parity", (d.get("parity") or {}).get("status"), (d.get("parity") or {}).get("bit_identical_to_world1"))
except Exception as e:
    print("ERR", e)
PY
