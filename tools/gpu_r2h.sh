#!/usr/bin/env bash
# Round-2 GPU session H (8 GPUs): smoke (tight timeout), bench --gpus 8, BASELINE configs[3] (4M disk fp64, energy tracking).
set -uo pipefail
O=gpurun_out/r2h; mkdir -p $O
W=${NB_WORLD:-8}
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29531 \
    tools/sharded_smoke.py > $O/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; grep "SMOKE\|Error" $O/smoke.log | tail -6
if [ $rc -ne 0 ]; then echo "smoke failed: stopping"; tail -20 $O/smoke.log; exit 0; fi
NB_BENCH_WATCHDOG_S=250 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29532 \
    bench.py --gpus $W --steps 10 --warmup 3 > $O/bench_n$W.json 2> $O/bench_n$W.err; echo "bench rc=$?"; tail -c 500 $O/bench_n$W.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29534 \
    tools/time_sharded.py 1048576 8 > $O/overlap_ab_n$W.log 2>&1; echo "ab rc=$?"; grep "world=\|Error" $O/overlap_ab_n$W.log | tail -8
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29533 \
    tools/run_configs.py c4 --ticks 3 > $O/config_c4.json 2> $O/config_c4.err; echo "c4 rc=$?"; tail -c 300 $O/config_c4.err; grep '"config"' $O/config_c4.json | cut -c1-900
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2h/bench_n$W.json").read().strip().splitlines()[-1])
    print("value %.4e ms/step %.3f e2e %.4e kernel_ms %.3f share %.5f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["kernel_share_of_step"]))
    print("  parity", d.get("parity"))
    for k,v in (d.get("lines") or {}).items(): print("  ", k, "%.4e" % v["value"], "ms %.2f" % v["ms_per_step"])
except Exception as e:
    print("ERR", e)
PY
