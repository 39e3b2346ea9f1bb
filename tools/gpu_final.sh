#!/bin/bash
# Round-end single-GPU validation + refreshed profiles, ordered by priority, every leg under its own timeout:
#   gpurun --timeout 840 -- 'bash tools/gpu_final.sh'
# 1. GPU parity tests  2. default bench line  3. ncu launch list of the bench step
# 4. ncu --set full of the five kernels of one forward SpMM (+ raw CSV page)  5. ncu launch list of a Pubmed
# train step  6. ncu --set full of the fused decoder's five kernels (Pubmed shape)  7. smoke()
mkdir -p gpurun_out
T0=$(date +%s)
timeout 540 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/bench.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spmm_ -c 60 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-pubmed --no-e2e > gpurun_out/ncu_list.log 2>&1; echo "rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/ncu_list.log
timeout 240 ncu --set full --clock-control none --import-source on -k regex:spmm_ --launch-skip 10 --launch-count 5 -f -o gpurun_out/spmm_full \
    python tools/spmm_sweep.py --iters 1 > gpurun_out/ncu_full.log 2>&1; echo "rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/ncu_full.log
ncu -i gpurun_out/spmm_full.ncu-rep --page raw --csv > gpurun_out/spmm_full_raw.csv 2>> gpurun_out/ncu_full.log
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_pubmed_step.csv \
    python tools/step_profile.py --steps 1 --which pubmed > gpurun_out/ncu_list_pubmed.log 2>&1; echo "rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/ncu_list_pubmed.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:dec_ --launch-skip 10 --launch-count 5 -f -o gpurun_out/dec_dense_default_full \
    python tools/dec_run.py pubmed 4 > gpurun_out/ncu_dec_default.log 2>&1; echo "rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/ncu_dec_default.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/smoke.log
tail -n 25 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/bench.err; cut -c1-1500 gpurun_out/bench.json; tail -n 2 gpurun_out/ncu_list.log gpurun_out/ncu_full.log gpurun_out/ncu_list_pubmed.log gpurun_out/ncu_dec_default.log gpurun_out/smoke.log
