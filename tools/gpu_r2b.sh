#!/usr/bin/env bash
# Round-2 GPU session B (2 GPUs): GPU tests, bench N=1 and N=2 (parity vs world 1 in-run), multi-GPU correctness tool.
set -uo pipefail
O=gpurun_out/r2b; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > $O/gputests.log 2>&1; echo "pytest rc=$?"; tail -25 $O/gputests.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench1 rc=$?"; tail -c 600 $O/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tools/run_sharded_check.py > $O/shard_check.log 2>&1; echo "shard check rc=$?"; grep -v "^\[" $O/shard_check.log | tail -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench2 rc=$?"; tail -c 1500 $O/bench_n2.err
NB_B200_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --steps 5 --warmup 3 --no-extras > $O/bench_n2_nooverlap.json 2> $O/bench_n2_nooverlap.err; echo "bench2 no-overlap rc=$?"
python - <<'PY'
import json
for f in ("bench_n1","bench_n2","bench_n2_nooverlap"):
    try:
        d=json.loads(open(f"gpurun_out/r2b/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.4e ms/step %.3f e2e %.4e kernel_ms %.3f share %.5f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["kernel_share_of_step"]))
        print("  parity", d.get("parity"))
        for k,v in (d.get("lines") or {}).items(): print("  ", k, "%.4e" % v["value"], "ms %.2f" % v["ms_per_step"])
    except Exception as e:
        print(f, "ERR", e)
PY
