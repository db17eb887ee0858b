#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cavpot_gpu.py -x -q -m gpu 2>&1 | tail -4
python -c "
from phaseshifts_b200 import api
print('selftest_cavpot mismatches:', api.selftest_cavpot(n=1<<24))
"
rm -f gpurun_out/r2_cavpot_step14.txt
for v in "" scripts/exp/libvar_u2m6.so scripts/exp/libvar_u2m5.so scripts/exp/libvar_u2m4.so scripts/exp/libvar_u1m8.so; do
  PHSH_B200_LIB=${v:+$PWD/$v} timeout 300 python scripts/exp/cavpot_dev_time.py 4096 2>&1 | tail -1 | tee -a gpurun_out/r2_cavpot_step14.txt
done
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:cavpot_mtz_sum -s 1 -c 1 python scripts/exp/cavpot_dev_time.py 4096 2 2>&1 | grep -E "gpu__time|fp64"
PHSH_B200_LIB=$PWD/scripts/exp/libvar_u2m5.so timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:cavpot_mtz_sum -s 1 -c 1 python scripts/exp/cavpot_dev_time.py 4096 2 2>&1 | grep -E "gpu__time|fp64"
