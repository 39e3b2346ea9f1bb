#!/bin/bash
# SpMM tuning sweep on the C4 workload:  gpurun --timeout 1800 -- 'bash tools/gpu_sweep.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "spmm" > gpurun_out/pytest_spmm.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_spmm.log
timeout 900 python tools/spmm_sweep.py --sweep --iters 5 --seg-lens 256,512 --blocks 64 --caches 0 --variants 0 --unrolls 4 --bins 1 --out gpurun_out/sweep8.json > gpurun_out/sweep8.log 2>&1
tail -n 3 gpurun_out/pytest_spmm.log; grep -v "^graph" gpurun_out/sweep8.log
