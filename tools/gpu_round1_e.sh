#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "spmm" > gpurun_out/pytest_spmm.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_spmm.log
timeout 900 python tools/spmm_sweep.py --sweep --iters 5 --seg-lens 64,128,256 --blocks 32,64,128 --caches 0 --variants 0 --out gpurun_out/sweep4.json > gpurun_out/sweep4.log 2>&1
tail -n 4 gpurun_out/pytest_spmm.log; sort -t: -k8 gpurun_out/sweep4.log | tail -3; grep BEST gpurun_out/sweep4.log
