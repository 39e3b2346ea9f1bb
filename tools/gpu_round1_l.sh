#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 --exchange push > gpurun_out/scale_8_push.log 2>&1; echo "rc=$?" >> gpurun_out/scale_8_push.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 20 --warmup 3 --exchange push --no-e2e > gpurun_out/scale_4_push.log 2>&1; echo "rc=$?" >> gpurun_out/scale_4_push.log
for f in gpurun_out/scale_8_push.log gpurun_out/scale_4_push.log; do grep -o '"ms_per_step": [0-9.]*' $f | head -1; grep -o '"fwd_ms": [0-9.]*' $f; done
