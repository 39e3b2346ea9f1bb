#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:'spmm_' -s 10 -c 5 --csv --log-file gpurun_out/launches_binned.csv python tools/spmm_sweep.py --iters 2 --seg-lens 512 > gpurun_out/ncu_binned.log 2>&1
tail -2 gpurun_out/ncu_binned.log
