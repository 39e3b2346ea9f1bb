#!/bin/bash
# Second GPU pass: bench line, launch list, full ncu capture of the SpMM kernels, wider sweep.
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 900 python tools/spmm_sweep.py --sweep --iters 5 --seg-lens 128,256,512,1024 --blocks 32,64,128 --caches 0 --out gpurun_out/sweep2.json > gpurun_out/sweep2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmm_' -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-pubmed --no-e2e > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'spmm_' -s 9 -c 3 -o gpurun_out/prof_spmm -f python tools/spmm_sweep.py --iters 2 > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/bench.log gpurun_out/sweep2.log gpurun_out/ncu_launches.log gpurun_out/ncu_full.log
