#!/usr/bin/env bash
# Round-2 GPU session E (2 GPUs): GPU tests, halo debug, bench N=1/N=2 with the three-window sharded tick, multi-GPU checks.
set -uo pipefail
O=gpurun_out/r2e; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > $O/gputests.log 2>&1; echo "pytest rc=$?"; tail -8 $O/gputests.log
timeout 200 python tools/debug_halo.py > $O/debug_halo.log 2>&1; cat $O/debug_halo.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench1 rc=$?"; tail -c 600 $O/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tools/run_sharded_check.py > $O/shard_check.log 2>&1; echo "shard check rc=$?"; grep "world=" $O/shard_check.log | tail -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 \
    tools/run_replicated_check.py > $O/replicated_check.log 2>&1; echo "replicated check rc=$?"; grep -i "replicated\|reference-style\|Error\|error" $O/replicated_check.log | tail -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench2 rc=$?"; tail -c 1500 $O/bench_n2.err
python - <<'PY'
import json
for f in ("bench_n1","bench_n2"):
    try:
        d=json.loads(open(f"gpurun_out/r2e/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.4e ms/step %.3f e2e %.4e kernel_ms %.3f share %.5f frac %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["kernel_share_of_step"], d["roofline"]["frac"]))
        print("  parity", d.get("parity"))
        for k,v in (d.get("lines") or {}).items(): print("  ", k, "%.4e" % v["value"], "ms %.2f" % v["ms_per_step"])
        if "small_n" in d: print("  small_n", d["small_n"])
    except Exception as e:
        print(f, "ERR", e)
PY
