#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "not full_size" > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
python tools/step_profile.py --steps 50 --which pubmed,cora > gpurun_out/sp_gemm.log 2>&1
tail -n 5 gpurun_out/pytest_gpu.log; grep "ms/step" gpurun_out/sp_gemm.log
