#!/bin/bash
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_cavpot_split_launches.csv python scripts/exp/cavpot_dev_time.py 4096 2 > gpurun_out/ncu_cavpot_launch.log 2>&1
grep -E "cavpot_" gpurun_out/r2_cavpot_split_launches.csv | tail -7 | cut -d, -f5,15-
