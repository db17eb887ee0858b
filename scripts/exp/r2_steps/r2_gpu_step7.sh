#!/bin/bash
# split cavpot: parity tests, device timing (with occupancy variants), per-kernel launch list; energy-axis kernels alone
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cavpot_gpu.py tests/test_boundary_gpu.py tests/test_full_size_gpu.py tests/test_rel_gpu.py tests/test_cav_wil_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_step7_tests.txt
cat gpurun_out/r2_step7_tests.txt
timeout 120 python scripts/exp/axis_time.py 2>&1 | tee gpurun_out/r2_axis_time.txt
for v in "" scripts/exp/libvar_prep2.so scripts/exp/libvar_mtz8.so scripts/exp/libvar_mtz4.so; do
  for S in 4096 1; do
    PHSH_B200_LIB=${v:+$PWD/$v} timeout 300 python scripts/exp/cavpot_dev_time.py $S 2>&1 | tail -1 | tee -a gpurun_out/r2_cavpot_split_variants.txt
  done
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_cavpot_split_launches.csv python scripts/exp/cavpot_dev_time.py 4096 2 > gpurun_out/ncu_cavpot_launch.log 2>&1
grep -E "cavpot_" gpurun_out/r2_cavpot_split_launches.csv | tail -12 | cut -d, -f5,15- 
