#!/bin/bash
# Full single-GPU validation as the driver runs it: smoke, GPU parity tests, both bench arms.
#   gpurun --timeout 2400 -- 'bash tools/gpu_validate.sh'
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref.log
tail -n 3 gpurun_out/smoke.log gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/bench.log gpurun_out/bench_ref.log
