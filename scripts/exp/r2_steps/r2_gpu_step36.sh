#!/bin/bash
mkdir -p gpurun_out
PHSH_B200_TRACE=1 timeout 120 python scripts/exp/hartfock_one.py x 2>&1 | grep -v "cluster size\|hartfock_file\|^\[phsh_b200\]   " > gpurun_out/r2_hartfock_times.txt
PHSH_B200_HARTFOCK_TEAM=1 timeout 120 python scripts/exp/hartfock_one.py x 2>&1 | grep -v "^\[phsh_b200\]" > gpurun_out/r2_hartfock_times_plain.txt
cat gpurun_out/r2_hartfock_times.txt | cut -c1-160
echo == plain; cat gpurun_out/r2_hartfock_times_plain.txt
