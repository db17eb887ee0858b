#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hartfock_gpu.py -x -q -m gpu 2>&1 | tail -60 | tee gpurun_out/step34_pytest.txt
