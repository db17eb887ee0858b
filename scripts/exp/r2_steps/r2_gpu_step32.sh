#!/bin/bash
# quick: hartfock parity, then the Re atom's timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hartfock_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/step32_pytest.txt
timeout 300 python scripts/exp/hartfock_one.py 2>&1 | tail -3 | tee gpurun_out/step32_times.txt
