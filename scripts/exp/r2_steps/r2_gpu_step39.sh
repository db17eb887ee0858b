#!/bin/bash
# cluster of 16 CTAs per atom (non-portable size): parity and timing against 8
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hartfock_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/step39_pytest.txt
for c in 16 8; do
echo "== cluster $c"
PHSH_B200_HARTFOCK_CLUSTER=$c PHSH_B200_TRACE=1 timeout 300 python scripts/exp/hartfock_one.py 2>&1 | grep -v "cluster size" | tail -3
done | tee gpurun_out/step39_times.txt
PHSH_B200_TRACE=1 timeout 300 python scripts/exp/hartfock_one.py 2>&1 | grep "cluster size\|Re batch" | tail -6 | tee -a gpurun_out/step39_times.txt
