#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu --no-pubmed --steps 20 > gpurun_out/bench_e2e.log 2>&1; echo "rc=$?" >> gpurun_out/bench_e2e.log
grep -o '"e2e": {[^}]*}' gpurun_out/bench_e2e.log; tail -n 1 gpurun_out/bench_e2e.log; grep -i "error\|Traceback" gpurun_out/bench_e2e.log | head -3
