#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_hartfock_gpu.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/step46_pytest.txt
