#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "cav or c2 or full_size or files or boundary or selftest" 2>&1 | tail -15 > gpurun_out/r2_gputests3.txt
cat gpurun_out/r2_gputests3.txt
python scripts/exp/cav_scale_probe.py > gpurun_out/r2_cav_scale2.txt 2>&1; cat gpurun_out/r2_cav_scale2.txt
python scripts/exp/cav_one_launch.py > gpurun_out/r2_cav_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cav_ps_kernel -s 1 -c 1 -o gpurun_out/r2_prof_cav2 python scripts/exp/cav_one_launch.py > gpurun_out/r2_ncu_cav.log 2>&1
tail -3 gpurun_out/r2_ncu_cav.log | cut -c1-200
