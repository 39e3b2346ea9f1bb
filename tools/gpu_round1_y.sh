#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench2_final.log 2>&1; echo "rc=$?" >> gpurun_out/bench2_final.log
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench2_ref.log 2>&1; echo "rc=$?" >> gpurun_out/bench2_ref.log
timeout 600 $TR tools/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check.log
grep -E "^\{|rc=" gpurun_out/bench2_final.log | cut -c1-400; grep -E "^\{|rc=" gpurun_out/bench2_ref.log | cut -c1-300; grep -E "DIST|rc=" gpurun_out/dist_check.log
