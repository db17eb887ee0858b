#!/bin/bash
mkdir -p gpurun_out
for c in 1 8; do echo "cluster=$c"; PHSH_B200_HARTFOCK_CLUSTER=$c timeout 600 python -m pytest tests/test_hartfock_gpu.py tests/test_boundary_gpu.py -x -q -m gpu 2>&1 | tail -4; done
for c in 1 2 4 8; do echo "cluster=$c"; PHSH_B200_HARTFOCK_CLUSTER=$c timeout 300 python scripts/exp/hartfock_one.py | tail -1; done
timeout 300 python scripts/exp/hartfock_one.py batches
timeout 300 python scripts/bench_paths.py 2>/dev/null | grep C7 | cut -c1-400
