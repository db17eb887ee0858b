#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cav_wil_gpu.py tests/test_full_size_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python scripts/exp/cav_scale_probe.py 2>&1 | tee gpurun_out/r2_cav_scale_unroll2.txt
ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread --clock-control none -k regex:cav_ps_kernel -s 1 -c 1 python scripts/exp/cav_one_launch.py 2>&1 | grep -E "gpu__time|fp64|inst_executed|registers" 
