#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o /tmp/block_probe scripts/exp/block_probe.cu && timeout 60 /tmp/block_probe | tee gpurun_out/r2_block_probe.txt
