#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cavpot_gpu.py tests/test_boundary_gpu.py -x -q -m gpu 2>&1 | tail -4
python -c "
from phaseshifts_b200 import api
print('selftest_cavpot mismatches:', api.selftest_cavpot(n=1<<24))
"
for S in 4096 1; do timeout 300 python scripts/exp/cavpot_dev_time.py $S 2>&1 | tail -1; done
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"cavpot_(mtz_sum|sumax)" -s 2 -c 2 python scripts/exp/cavpot_dev_time.py 4096 2 2>&1 | grep -E "cavpot_|gpu__time|fp64|inst_executed|issue_active"
