#!/usr/bin/env bash
# Session W (2 GPUs): final binary — sharded smoke + shard check + bench --gpus 2.
set -uo pipefail
O=gpurun_out/r2w; mkdir -p $O
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 \
    tools/sharded_smoke.py > $O/smoke.log 2>&1; echo "smoke rc=$?"; grep "SMOKE\|Error" $O/smoke.log | tail -n 3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29582 \
    tools/run_sharded_check.py > $O/shard_check.log 2>&1; echo "shard check rc=$?"; grep -c " OK" $O/shard_check.log; grep "MISMATCH\|Error" $O/shard_check.log | head -n 5
NB_BENCH_WATCHDOG_S=200 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29583 \
    bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2w/bench_n2.json").read().strip().splitlines()[-1])
print("value %.4e ms/step %.3f e2e %.4e parity %s %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"]["status"], d["parity"]["bit_identical_to_world1"]))
PY
