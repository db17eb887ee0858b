#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cavpot_gpu.py tests/test_boundary_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_step11_tests.txt
cat gpurun_out/r2_step11_tests.txt
for S in 4096 1; do timeout 300 python scripts/exp/cavpot_dev_time.py $S 2>&1 | tail -1; done
timeout 300 python scripts/exp/latency_breakdown.py 2>&1 | tee gpurun_out/r2_latency_breakdown.txt
timeout 300 python scripts/exp/chain_latency.py 2>&1 | tee gpurun_out/r2_chain_latency.json
