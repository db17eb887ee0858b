#!/bin/bash
# team sweeps of hartfock: parity (all modes), then timing of the Re atom with and without teams
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hartfock_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/step29_pytest.txt
cat gpurun_out/step29_pytest.txt
for m in 0 1; do
  echo "== PHSH_B200_HARTFOCK_TEAM=$m"
  PHSH_B200_HARTFOCK_TEAM=$m timeout 300 python scripts/exp/hartfock_one.py x 2>&1 | tail -9
done | tee gpurun_out/step29_times.txt
