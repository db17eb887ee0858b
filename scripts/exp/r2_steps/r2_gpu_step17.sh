#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cavpot_gpu.py tests/test_boundary_gpu.py tests/test_guard_gpu.py -x -q -m gpu 2>&1 | tail -3
PHSH_B200_TRACE=1 timeout 300 python scripts/exp/cavpot_time.py 4096 2>&1 | tail -16
