#!/bin/bash
mkdir -p gpurun_out
python tools/step_profile.py --steps 50 --which pubmed,zinc > gpurun_out/sp_r2.log 2>&1
python tools/step_profile.py --steps 50 --which pubmed,zinc --tune dec_rows=1 > gpurun_out/sp_r1.log 2>&1
grep "ms/step" gpurun_out/sp_r2.log gpurun_out/sp_r1.log
