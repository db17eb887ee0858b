#!/bin/bash
# phase counters of the team sweeps (library built with -DPHSH_HF_PROFILE), then parity
mkdir -p gpurun_out
PHSH_B200_TRACE=1 timeout 300 python scripts/exp/hartfock_one.py 2>&1 | grep -v "cluster size\|cluster of" | tail -8 | tee gpurun_out/step30_profile.txt
timeout 900 python -m pytest tests/test_hartfock_gpu.py -x -q -m gpu 2>&1 | grep -v "^hf profile" | tail -15 | tee gpurun_out/step30_pytest.txt
