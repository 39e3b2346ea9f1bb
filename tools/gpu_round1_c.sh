#!/bin/bash
# 2-GPU pass: distributed parity check, bench with both exchange mechanisms, 1-GPU bench with the new defaults.
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check.log
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 --exchange nccl > gpurun_out/bench2_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/bench2_nccl.log
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 --exchange p2p > gpurun_out/bench2_p2p.log 2>&1; echo "rc=$?" >> gpurun_out/bench2_p2p.log
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu --no-pubmed --no-e2e > gpurun_out/bench1.log 2>&1
tail -n 4 gpurun_out/dist_check.log gpurun_out/bench2_nccl.log gpurun_out/bench2_p2p.log gpurun_out/bench1.log
