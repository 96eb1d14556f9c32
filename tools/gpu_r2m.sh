#!/usr/bin/env bash
# Session M (2 GPUs): gather timing + arrangement A/B, final bench --gpus 2 (both arms).
set -uo pipefail
O=gpurun_out/r2m; mkdir -p $O
W=${NB_WORLD:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29551 \
    tools/time_sharded.py > $O/overlap_ab_n$W.log 2>&1; echo "ab rc=$?"; grep "world=\|Error" $O/overlap_ab_n$W.log | tail -12
NB_BENCH_WATCHDOG_S=250 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29552 \
    bench.py --gpus $W --steps 5 --warmup 3 > $O/bench_n$W.json 2> $O/bench_n$W.err; echo "bench rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29553 \
    bench.py --impl reference --gpus $W --steps 2 --warmup 1 > $O/bench_reference_n$W.json 2> $O/bench_reference_n$W.err; echo "reference arm rc=$?"; cut -c1-400 $O/bench_reference_n$W.json
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
