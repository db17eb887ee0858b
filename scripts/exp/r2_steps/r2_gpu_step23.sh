#!/bin/bash
timeout 600 python -m pytest tests/test_hartfock_gpu.py tests/test_boundary_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python scripts/exp/hartfock_one.py batches | tail -7
PHSH_B200_HARTFOCK_CLUSTER=1 timeout 300 python scripts/exp/hartfock_one.py | tail -1
timeout 300 python scripts/bench_paths.py 2>/dev/null | grep C7 | cut -c1-500
