#!/bin/bash
# last look: smoke and the whole GPU suite on the final tree
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/final_smoke.txt
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/final_pytest_gpu.txt
