#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "stream or spmm" > gpurun_out/pytest_stream.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_stream.log
timeout 900 python tools/spmm_sweep.py --sweep --iters 5 --seg-lens 128,512,100000000 --variants 1,2 --stages 2,3,4 --out gpurun_out/sweep3.json > gpurun_out/sweep3.log 2>&1
tail -n 15 gpurun_out/pytest_stream.log; cat gpurun_out/sweep3.log
