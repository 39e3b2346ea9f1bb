#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dec_dense' -s 3 -c 1 -o gpurun_out/prof_dec -f python tools/step_profile.py --steps 1 --which pubmed > gpurun_out/ncu_dec.log 2>&1
tail -3 gpurun_out/ncu_dec.log
