#!/bin/bash
# the two-GPU test on the final tree (gpurun --gpus 2)
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_distributed_nccl_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/step47_two_gpu.txt
