#!/usr/bin/env bash
# Round-2 GPU session G (2 GPUs): sharded smoke first (tight timeouts everywhere), then the multi-GPU check and bench N=2.
set -uo pipefail
O=gpurun_out/r2i; mkdir -p $O
W=${NB_WORLD:-2}
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29521 \
    tools/sharded_smoke.py > $O/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; grep "rank\|SMOKE\|Error" $O/smoke.log | tail -12
if [ $rc -ne 0 ]; then
  echo "windowed path failed: retrying the smoke with NB_B200_OVERLAP=0"
  NB_B200_OVERLAP=0 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29522 \
      tools/sharded_smoke.py > $O/smoke_nooverlap.log 2>&1; echo "smoke (no overlap) rc=$?"; grep "rank\|SMOKE\|Error" $O/smoke_nooverlap.log | tail -12
  exit 0
fi
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29523 \
    tools/run_sharded_check.py > $O/shard_check.log 2>&1; echo "shard check rc=$?"; grep "world=" $O/shard_check.log | tail -20
NB_BENCH_WATCHDOG_S=250 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29524 \
    bench.py --gpus $W --steps 5 --warmup 3 > $O/bench_n$W.json 2> $O/bench_n$W.err; echo "bench rc=$?"; tail -c 600 $O/bench_n$W.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2i/bench_n$W.json").read().strip().splitlines()[-1])
    print("value %.4e ms/step %.3f e2e %.4e kernel_ms %.3f share %.5f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["kernel_share_of_step"]))
    print("  parity", d.get("parity"))
    for k,v in (d.get("lines") or {}).items(): print("  ", k, "%.4e" % v["value"], "ms %.2f" % v["ms_per_step"])
except Exception as e:
    print("ERR", e)
PY
