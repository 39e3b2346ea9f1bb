#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python tools/spmm_sweep.py --sweep --iters 5 --seg-lens 256,512 --blocks 64,128 --caches 0 --variants 0 --unrolls 2,4 --out gpurun_out/sweep5.json > gpurun_out/sweep5.log 2>&1
grep BEST gpurun_out/sweep5.log
