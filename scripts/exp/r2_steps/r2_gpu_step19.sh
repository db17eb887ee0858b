#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hartfock_gpu.py -x -q -m gpu 2>&1 | tail -8
timeout 300 python scripts/exp/hartfock_one.py
