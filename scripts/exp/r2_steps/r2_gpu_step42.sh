#!/bin/bash
# terms dealt to the threads sorted by kind: parity, then the batch of 148 Re atoms in both modes
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_hartfock_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 120 python scripts/exp/hartfock_batch_ab.py 148 2>&1 | tee gpurun_out/step42_ab.txt
