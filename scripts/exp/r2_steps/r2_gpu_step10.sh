#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cavpot_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_step10_tests.txt
cat gpurun_out/r2_step10_tests.txt
rm -f gpurun_out/r2_cavpot_step10.txt
for v in "" scripts/exp/libvar_u1m6.so scripts/exp/libvar_u2m5.so scripts/exp/libvar_u2m4.so scripts/exp/libvar_u4m4.so scripts/exp/libvar_prep2.so scripts/exp/libvar_prep3.so; do
  PHSH_B200_LIB=${v:+$PWD/$v} timeout 300 python scripts/exp/cavpot_dev_time.py 4096 2>&1 | tail -1 | tee -a gpurun_out/r2_cavpot_step10.txt
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_cavpot_split_launches.csv python scripts/exp/cavpot_dev_time.py 4096 2 > gpurun_out/ncu_cavpot_launch.log 2>&1
grep -E "cavpot_" gpurun_out/r2_cavpot_split_launches.csv | tail -6 | cut -d, -f5,15- 
