#!/usr/bin/env bash
# Session R (2 GPUs): the replicated mode of GalaxySimulation after it became opt-in, incl. the unchanged driver through run_script.
set -uo pipefail
O=gpurun_out/r2r; mkdir -p $O
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
    tools/run_replicated_check.py > $O/replicated_check.log 2>&1; echo "replicated check rc=$?"; grep -i "replicated\|reference-style\|Error\|error\|OK\|MISMATCH" $O/replicated_check.log | tail -n 20
