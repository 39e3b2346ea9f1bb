#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "packed or trainers" > gpurun_out/pytest_packed.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_packed.log
python tools/collate_bench.py > gpurun_out/collate_bench.log 2>&1
tail -n 12 gpurun_out/pytest_packed.log; cat gpurun_out/collate_bench.log | grep -v Dataset
