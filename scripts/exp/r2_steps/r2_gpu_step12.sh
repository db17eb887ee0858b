#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cavpot_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_step12_tests.txt
cat gpurun_out/r2_step12_tests.txt
timeout 300 python scripts/exp/hartfock_one.py
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hartfock -c 1 -f -o gpurun_out/r2_prof_hartfock python scripts/exp/hartfock_one.py > gpurun_out/ncu_hartfock.log 2>&1
tail -3 gpurun_out/ncu_hartfock.log
