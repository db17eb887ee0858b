#!/bin/bash
mkdir -p gpurun_out
./bin_tmp/lat_probe | tee gpurun_out/r2_latency_probe.txt
