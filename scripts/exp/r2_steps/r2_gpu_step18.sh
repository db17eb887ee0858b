#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cav_wil_gpu.py tests/test_full_size_gpu.py tests/test_files_gpu.py -x -q -m gpu 2>&1 | tail -3
for cg in 4 2 1 0; do echo "PHSH_B200_WIL_CGROUPS=$cg"; PHSH_B200_WIL_CGROUPS=$cg timeout 300 python scripts/exp/wil_time.py 2>&1 | tail -1; done
