#!/bin/bash
# 8-GPU box: parity check at 8 ranks, then the scaling series 1/2/4/8 as the driver runs it.
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
run() { N=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 "$@"; }
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 tools/dist_check.py > gpurun_out/dist_check8.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check8.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu --no-pubmed > gpurun_out/scale_1.log 2>&1
for N in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_$N.log 2>&1; echo "rc=$?" >> gpurun_out/scale_$N.log
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 --exchange nccl --no-e2e > gpurun_out/scale_8_nccl.log 2>&1
tail -n 3 gpurun_out/dist_check8.log
for f in gpurun_out/scale_*.log; do echo $f; grep -o '"value": [0-9.e+]*, "unit": "edges/s", "n_gpus": [0-9]*\|"ms_per_step": [0-9.]*\|"halo_rows_rank0": [0-9]*\|"fwd_ms": [0-9.]*' $f | head -5 | tr '\n' ' '; echo; done
