#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cavpot_gpu.py tests/test_boundary_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_step9_tests.txt
cat gpurun_out/r2_step9_tests.txt
python -c "
from phaseshifts_b200 import api
print('selftest_cavpot mismatches:', api.selftest_cavpot(n=1<<24))
"
for S in 4096 1; do timeout 300 python scripts/exp/cavpot_dev_time.py $S 2>&1 | tail -1 | tee -a gpurun_out/r2_cavpot_step9.txt; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_cavpot_split_launches.csv python scripts/exp/cavpot_dev_time.py 4096 2 > gpurun_out/ncu_cavpot_launch.log 2>&1
grep -E "cavpot_" gpurun_out/r2_cavpot_split_launches.csv | tail -6 | cut -d, -f5,15- 
