#!/bin/bash
mkdir -p gpurun_out
PHSH_B200_TRACE=1 timeout 60 python scripts/exp/hartfock_one.py 2>&1 | grep -v "cluster size\|cluster of" | tail -5 | tee gpurun_out/step31_profile.txt
