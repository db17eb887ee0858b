#!/bin/bash
# dependent-issue latencies of one warp (scripts/exp/lat_probe.cu), compiled on the box
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o /tmp/lat_probe scripts/exp/lat_probe.cu && /tmp/lat_probe | tee gpurun_out/r2_latency_probe.txt
