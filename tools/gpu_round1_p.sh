#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmm_' -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-pubmed --no-e2e > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'spmm_' -s 10 -c 5 -o gpurun_out/prof_spmm_binned -f python tools/spmm_sweep.py --iters 2 --seg-lens 512 > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log; tail -n 2 gpurun_out/bench.log gpurun_out/bench_ref.log
