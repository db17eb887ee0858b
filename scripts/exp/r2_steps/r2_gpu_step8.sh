#!/bin/bash
# energy-axis kernels after the constant-map fast path; ncu --set full of the split cavpot kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rel_gpu.py tests/test_files_gpu.py tests/test_full_size_gpu.py tests/test_conphas_files_gpu.py tests/test_cav_wil_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_step8_tests.txt
cat gpurun_out/r2_step8_tests.txt
timeout 120 python scripts/exp/axis_time.py 2>&1 | tee gpurun_out/r2_axis_time.txt
CMD="python scripts/exp/cavpot_dev_time.py 4096 2"
for k in mtz_sum sumax prep classify; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:cavpot_${k} -s 1 -c 1 -o gpurun_out/r2_prof_cavpot_$k -f $CMD > gpurun_out/ncu_cavpot_$k.log 2>&1
  tail -2 gpurun_out/ncu_cavpot_$k.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
