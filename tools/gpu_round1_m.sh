#!/bin/bash
set -x
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tools/scale_compare.py --modes push,nccl --steps 10 > gpurun_out/scale_compare_$N.log 2>&1; echo "rc=$?" >> gpurun_out/scale_compare_$N.log
grep -E "SCALE_COMPARE|rc=|rror" gpurun_out/scale_compare_$N.log
