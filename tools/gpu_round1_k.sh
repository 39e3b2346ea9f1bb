#!/bin/bash
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check.log
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 --exchange push --no-e2e > gpurun_out/bench2_push.log 2>&1; echo "rc=$?" >> gpurun_out/bench2_push.log
grep -E "exchange=|DIST|rc=" gpurun_out/dist_check.log; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench2_push.log; grep -i "error|Traceback" gpurun_out/bench2_push.log | head
