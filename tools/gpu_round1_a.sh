#!/bin/bash
# First GPU pass: smoke, parity tests, tuning sweep, short bench.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python tools/spmm_sweep.py --sweep --iters 5 --out gpurun_out/sweep.json > gpurun_out/sweep.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -5 gpurun_out/smoke.log gpurun_out/pytest_gpu.log gpurun_out/bench.log
tail -3 gpurun_out/sweep.log
