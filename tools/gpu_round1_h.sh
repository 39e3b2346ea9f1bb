#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
python tools/step_profile.py --steps 50 > gpurun_out/step_profile2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_pubmed2.csv python tools/step_profile.py --steps 1 --which pubmed > gpurun_out/ncu_pubmed2.log 2>&1
tail -n 5 gpurun_out/pytest_gpu.log; cat gpurun_out/step_profile2.log
