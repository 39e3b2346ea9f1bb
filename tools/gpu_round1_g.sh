#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/step_profile.py --steps 50 > gpurun_out/step_profile.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_pubmed.csv python tools/step_profile.py --steps 1 --which pubmed > gpurun_out/ncu_pubmed.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_zinc.csv python tools/step_profile.py --steps 1 --which zinc > gpurun_out/ncu_zinc.log 2>&1
cat gpurun_out/step_profile.log
