#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "decoder or train_step or graph_step or dense_reference" > gpurun_out/pytest_dec.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_dec.log
python tools/step_profile.py --steps 50 > gpurun_out/step_profile3.log 2>&1
tail -n 4 gpurun_out/pytest_dec.log; grep "ms/step" gpurun_out/step_profile3.log
