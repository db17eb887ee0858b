#!/bin/bash
# short-fused check of an experimental hartfock build: every command under a 60 s timeout
mkdir -p gpurun_out
timeout 60 python scripts/exp/hartfock_one.py 2>&1 | tail -3 | tee gpurun_out/step40_times.txt
timeout 90 python -m pytest tests/test_hartfock_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/step40_pytest.txt
