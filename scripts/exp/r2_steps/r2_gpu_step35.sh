#!/bin/bash
# after the hartfock restructuring: the whole GPU suite, the hartfock file three more times (the replay writes ranges
# concurrently: any race would show as a flaky bit-identity test), timings of single atoms and batches
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/step35_pytest_gpu.txt
for i in 1 2 3; do python -m pytest tests/test_hartfock_gpu.py -x -q -m gpu 2>&1 | tail -1; done | tee gpurun_out/step35_repeat.txt
PHSH_B200_TRACE=1 python scripts/exp/hartfock_one.py x 2>&1 | grep -v "cluster size" | tail -14 | tee gpurun_out/step35_times.txt
